// Device-resident camera algebra between the stages of a batched prediction (see glue_math.cuh): four tiny
// kernels turn detector output into the selector's crop jobs, detector + selector output into the initial poses,
// poses into the refiner's crop jobs + camera tensors, and the refiner's output into the next poses -- so that
// detect -> select -> refine x N is ONE stream-ordered sequence (one CUDA graph) with no host in the loop.
// The *_host entry points run the same functions on host memory (unit tests against geometry.py, no GPU needed).
#include "common.cuh"
#include "glue_math.cuh"

namespace g6d {
using namespace glue;

G6D_HD void do_detection_job(int i, const float* det_out, const uint8_t* frames, int rows, int cols, int size, g6d_warp_job* jobs) {
    g6d_warp_job j;
    j.src = frames + (long long)i * rows * cols * 3;
    j.rows = rows; j.cols = cols;
    detection_crop_matrix(det_out[i * 4], det_out[i * 4 + 1], det_out[i * 4 + 2], size, j.M);
    jobs[i] = j;
}

G6D_HD void do_initial_pose(int i, const float* det_out, const long long* sel_idx, const float* sel_out, const g6d_glue_refs& r,
                            const g6d_glue_camera* cams, double* poses) {
    const long long k = sel_idx[i];
    pose_from_similarity(det_out[i * 4], det_out[i * 4 + 1], det_out[i * 4 + 2], sel_out[i * 2], r.poses + k * 12, r.cen + k * 2,
                         r.f[k], r.dist[k], cams[i].Kinv, cams[i].f, cams[i].f_sq, r.center, poses + (long long)i * 12);
}

G6D_HD NormParams norm_of(const g6d_glue_views& v) {
    NormParams n;
    n.scale = v.norm_scale;
    n.offset[0] = v.norm_offset[0]; n.offset[1] = v.norm_offset[1]; n.offset[2] = v.norm_offset[2];
    return n;
}

// frame part of one refinement problem: normalise, look at the object, choose the ref_num nearest views
G6D_HD void do_refine_frame(int i, const g6d_glue_views& v, const g6d_glue_camera* cams, const double* poses, int in_f32,
                            FrameProblem& fp, int* chosen /* [ref_num] table rows */) {
    float pose_n[12];
    normalize_pose(poses + (long long)i * 12, in_f32, norm_of(v), pose_n);
    refine_frame(pose_n, cams[i].K, cams[i].Kinv, cams[i].f, v.size, v.size_scale, fp);
    // database_utils.py:125-139: the ref_num views of the FPS subset with the largest cosine to the viewing direction
    for (int r = 0; r < v.ref_num; ++r) {
        int best = -1; float bv = 0.f;
        for (int e = 0; e < v.n_even; ++e) {
            bool used = false;
            for (int q = 0; q < r; ++q) used |= chosen[q] == v.even_idx[e];
            if (used) continue;
            const float* d = v.even_dirs + e * 3;
            const float dot = (d[0] * fp.qdir[0] + d[1] * fp.qdir[1]) + d[2] * fp.qdir[2];
            if (best < 0 || dot > bv) { best = e; bv = dot; }
        }
        chosen[r] = v.even_idx[best];
    }
}

G6D_HD void store_frame(int i, const g6d_glue_views& v, const FrameProblem& fp, const uint8_t* frames, int rows, int cols,
                        g6d_warp_job* jobs, float* que_K, float* que_pose, float* rect) {
    for (int e = 0; e < 9; ++e) que_K[i * 9 + e] = fp.K_warp[e];
    for (int e = 0; e < 12; ++e) { que_pose[i * 12 + e] = fp.pose_warp[e]; rect[i * 12 + e] = fp.rect[e]; }
    g6d_warp_job j;
    j.src = frames + (long long)i * rows * cols * 3;
    j.rows = rows; j.cols = cols;
    inv3_cv(fp.que_H, j.M);
    jobs[(long long)i * (v.ref_num + 1)] = j;
}

G6D_HD void do_refine_view(int i, int r, int row, const g6d_glue_views& v, const double* Rq, g6d_warp_job* jobs, float* ref_Ks,
                           float* ref_poses, int* ref_rows) {
    ViewProblem vp;
    refine_view(Rq, v.poses + (long long)row * 12, v.R_look + (long long)row * 9, v.RlookR + (long long)row * 9, v.f[row],
                v.Kinv + (long long)row * 9, v.size, vp);
    const long long o = (long long)i * v.ref_num + r;
    for (int e = 0; e < 9; ++e) ref_Ks[o * 9 + e] = vp.K[e];
    for (int e = 0; e < 12; ++e) ref_poses[o * 12 + e] = vp.pose[e];
    ref_rows[o] = row;
    g6d_warp_job j;
    j.src = reinterpret_cast<const uint8_t*>(v.src[row]);
    j.rows = v.rows[row]; j.cols = v.cols[row];
    inv3_cv(vp.H, j.M);
    jobs[(long long)i * (v.ref_num + 1) + 1 + r] = j;
}

G6D_HD void do_apply(int i, const g6d_glue_views& v, const float* que_pose, const float* que_K, const float* rect, const float* net_out,
                     double* poses) {
    float p[12];
    apply_refinement(que_pose + i * 12, que_K + i * 9, rect + i * 12, net_out + i * 7, norm_of(v), p);
    for (int e = 0; e < 12; ++e) poses[(long long)i * 12 + e] = (double)p[e];
}

// ------------------------------------------------------------------------------------------ kernels
__global__ void glue_detection_jobs_kernel(const float* det_out, const uint8_t* frames, int rows, int cols, int qn, int size,
                                           g6d_warp_job* jobs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < qn) do_detection_job(i, det_out, frames, rows, cols, size, jobs);
}
__global__ void glue_initial_poses_kernel(const float* det_out, const long long* sel_idx, const float* sel_out, const g6d_glue_refs r,
                                          const g6d_glue_camera* cams, int qn, double* poses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < qn) do_initial_pose(i, det_out, sel_idx, sel_out, r, cams, poses);
}
// one block per frame: thread 0 does the frame part, then one thread per selected view
constexpr int kGlueMaxViews = 8;
__global__ void __launch_bounds__(32) glue_refine_problems_kernel(const g6d_glue_views v, const g6d_glue_camera* cams,
                                                                  const uint8_t* frames, int rows, int cols, const double* poses,
                                                                  int in_f32, g6d_warp_job* jobs, float* que_K, float* que_pose,
                                                                  float* rect, float* ref_Ks, float* ref_poses, int* ref_rows) {
    __shared__ double s_Rq[9];
    __shared__ int s_rows[kGlueMaxViews];
    const int i = blockIdx.x;
    if (threadIdx.x == 0) {
        FrameProblem fp;
        do_refine_frame(i, v, cams, poses, in_f32, fp, s_rows);
        store_frame(i, v, fp, frames, rows, cols, jobs, que_K, que_pose, rect);
        for (int e = 0; e < 9; ++e) s_Rq[e] = fp.Rq[e];
    }
    __syncthreads();
    if ((int)threadIdx.x < v.ref_num) do_refine_view(i, threadIdx.x, s_rows[threadIdx.x], v, s_Rq, jobs, ref_Ks, ref_poses, ref_rows);
}
__global__ void glue_apply_kernel(const g6d_glue_views v, const float* que_pose, const float* que_K, const float* rect,
                                  const float* net_out, int qn, double* poses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < qn) do_apply(i, v, que_pose, que_K, rect, net_out, poses);
}

static bool views_ok(const g6d_glue_views* v) {
    return v && v->poses && v->R_look && v->RlookR && v->f && v->Kinv && v->src && v->rows && v->cols && v->even_idx && v->even_dirs &&
           v->n_views > 0 && v->n_even > 0 && v->ref_num >= 2 && v->ref_num < kGlueMaxViews && v->ref_num <= v->n_even && v->size > 0 &&
           v->norm_scale > 0;
}

}  // namespace g6d

using namespace g6d;

extern "C" int g6d_glue_detection_jobs(const float* det_out, const uint8_t* frames, int rows, int cols, int qn, int size,
                                       g6d_warp_job* jobs, g6d_stream_t stream) {
    G6D_REQUIRE(det_out && frames && jobs && qn > 0 && rows > 0 && cols > 0 && size > 0, "g6d_glue_detection_jobs: bad args");
    glue_detection_jobs_kernel<<<ceil_div(qn, 32), 32, 0, as_stream(stream)>>>(det_out, frames, rows, cols, qn, size, jobs);
    G6D_CHECK_LAUNCH("g6d_glue_detection_jobs");
    return G6D_OK;
}
extern "C" int g6d_glue_detection_jobs_host(const float* det_out, const uint8_t* frames, int rows, int cols, int qn, int size,
                                            g6d_warp_job* jobs) {
    G6D_REQUIRE(det_out && jobs && qn > 0, "g6d_glue_detection_jobs_host: bad args");
    for (int i = 0; i < qn; ++i) do_detection_job(i, det_out, frames, rows, cols, size, jobs);
    return G6D_OK;
}

extern "C" int g6d_glue_initial_poses(const float* det_out, const long long* sel_idx, const float* sel_out, const g6d_glue_refs* refs,
                                      const g6d_glue_camera* cams, int qn, double* poses, g6d_stream_t stream) {
    G6D_REQUIRE(det_out && sel_idx && sel_out && refs && refs->poses && refs->cen && refs->f && refs->dist && cams && poses && qn > 0,
                "g6d_glue_initial_poses: bad args");
    glue_initial_poses_kernel<<<ceil_div(qn, 32), 32, 0, as_stream(stream)>>>(det_out, sel_idx, sel_out, *refs, cams, qn, poses);
    G6D_CHECK_LAUNCH("g6d_glue_initial_poses");
    return G6D_OK;
}
extern "C" int g6d_glue_initial_poses_host(const float* det_out, const long long* sel_idx, const float* sel_out, const g6d_glue_refs* refs,
                                           const g6d_glue_camera* cams, int qn, double* poses) {
    G6D_REQUIRE(det_out && sel_idx && sel_out && refs && cams && poses && qn > 0, "g6d_glue_initial_poses_host: bad args");
    for (int i = 0; i < qn; ++i) do_initial_pose(i, det_out, sel_idx, sel_out, *refs, cams, poses);
    return G6D_OK;
}

extern "C" int g6d_glue_refine_problems(const g6d_glue_views* views, const g6d_glue_camera* cams, const uint8_t* frames, int rows,
                                        int cols, const double* poses, int poses_are_f32, int qn, g6d_warp_job* jobs, float* que_K,
                                        float* que_pose, float* rect, float* ref_Ks, float* ref_poses, int* ref_rows,
                                        g6d_stream_t stream) {
    G6D_REQUIRE(views_ok(views), "g6d_glue_refine_problems: incomplete view tables (2 <= ref_num < %d)", kGlueMaxViews);
    G6D_REQUIRE(cams && frames && poses && jobs && que_K && que_pose && rect && ref_Ks && ref_poses && ref_rows && qn > 0 && rows > 0 &&
                    cols > 0, "g6d_glue_refine_problems: bad args");
    glue_refine_problems_kernel<<<qn, 32, 0, as_stream(stream)>>>(*views, cams, frames, rows, cols, poses, poses_are_f32, jobs, que_K,
                                                                  que_pose, rect, ref_Ks, ref_poses, ref_rows);
    G6D_CHECK_LAUNCH("g6d_glue_refine_problems");
    return G6D_OK;
}
extern "C" int g6d_glue_refine_problems_host(const g6d_glue_views* views, const g6d_glue_camera* cams, const uint8_t* frames, int rows,
                                             int cols, const double* poses, int poses_are_f32, int qn, g6d_warp_job* jobs, float* que_K,
                                             float* que_pose, float* rect, float* ref_Ks, float* ref_poses, int* ref_rows) {
    G6D_REQUIRE(views_ok(views), "g6d_glue_refine_problems_host: incomplete view tables");
    G6D_REQUIRE(cams && poses && jobs && que_K && que_pose && rect && ref_Ks && ref_poses && ref_rows && qn > 0, "g6d_glue_refine_problems_host: bad args");
    for (int i = 0; i < qn; ++i) {
        FrameProblem fp;
        int chosen[kGlueMaxViews];
        do_refine_frame(i, *views, cams, poses, poses_are_f32, fp, chosen);
        store_frame(i, *views, fp, frames, rows, cols, jobs, que_K, que_pose, rect);
        for (int r = 0; r < views->ref_num; ++r) do_refine_view(i, r, chosen[r], *views, fp.Rq, jobs, ref_Ks, ref_poses, ref_rows);
    }
    return G6D_OK;
}

extern "C" int g6d_glue_apply_refinements(const g6d_glue_views* views, const float* que_pose, const float* que_K, const float* rect,
                                          const float* net_out, int qn, double* poses, g6d_stream_t stream) {
    G6D_REQUIRE(views && views->norm_scale > 0 && que_pose && que_K && rect && net_out && poses && qn > 0, "g6d_glue_apply_refinements: bad args");
    glue_apply_kernel<<<ceil_div(qn, 32), 32, 0, as_stream(stream)>>>(*views, que_pose, que_K, rect, net_out, qn, poses);
    G6D_CHECK_LAUNCH("g6d_glue_apply_refinements");
    return G6D_OK;
}
extern "C" int g6d_glue_apply_refinements_host(const g6d_glue_views* views, const float* que_pose, const float* que_K, const float* rect,
                                               const float* net_out, int qn, double* poses) {
    G6D_REQUIRE(views && views->norm_scale > 0 && que_pose && que_K && rect && net_out && poses && qn > 0, "g6d_glue_apply_refinements_host: bad args");
    for (int i = 0; i < qn; ++i) do_apply(i, *views, que_pose, que_K, rect, net_out, poses);
    return G6D_OK;
}
