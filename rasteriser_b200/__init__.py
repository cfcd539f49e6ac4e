"""rasteriser_b200 -- B200-native frame path of the canmom/rasteriser renderer.

librast_b200.so (hand-written sm_100a CUDA behind the extern "C" rast_* ABI of include/rast.h) is
the product; this package is its thin host-side mirror of the reference's draw_frame interface."""
from .api import Args, Material, Renderer, RastError, draw_frame, frame_matrices, spin_angle  # noqa: F401
