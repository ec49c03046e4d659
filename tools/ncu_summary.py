"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into tracked files under profiles/:
    python tools/ncu_summary.py r02 s2=gpurun_out/ncu_s2.ncu-rep r2=gpurun_out/ncu_r2.ncu-rep conv=gpurun_out/ncu_conv.ncu-rep
writes profiles/ncu_<round>_<key>.csv (the metrics of the raw page that the roofline discussion uses) and
profiles/ncu_traffic.json ({key: dram__bytes_read.sum + dram__bytes_write.sum per launch}, read by bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct')
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rnd, traffic = sys.argv[1], {}
    tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for arg in sys.argv[2:]:
        key, path = arg.split('=')
        h, u, launches = raw(path)
        with open(os.path.join(ROOT, 'profiles', f'ncu_{rnd}_{key}.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['kernel', 'metric', 'unit', 'value'])
            for r in launches:
                name = r[h.index('Kernel Name')]
                for m in KEEP:
                    if m in h:
                        w.writerow([name, m, u[h.index(m)], r[h.index(m)]])
        r = launches[0]
        tot = sum(float(r[h.index(m)].replace(',', '')) * UNIT[u[h.index(m)]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        traffic[key] = tot
        print(key, r[h.index('Kernel Name')][:60], 'time', r[h.index('gpu__time_duration.sum')], u[h.index('gpu__time_duration.sum')],
              'dram bytes', tot)
    json.dump(traffic, open(tpath, 'w'), indent=1)


if __name__ == '__main__':
    main()
