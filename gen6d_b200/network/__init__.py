"""Drop-in replacement for the reference's `network` package (network/__init__.py:1-9):
`name2network[cfg['network']](cfg)` -> module with the reference's methods and checkpoint keys."""
from .detector import Detector
from .refiner import VolumeRefiner
from .selector import ViewpointSelector

name2network = {
    'refiner': VolumeRefiner,
    'detector': Detector,
    'selector': ViewpointSelector,
}
