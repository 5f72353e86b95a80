"""voidin_b200 — B200-native (sm_100a CUDA) replacement for the acceleration-structure hot path of pudnax/voidin:
`crates/bvh` BLAS build, TLAS build and closest-hit / any-hit traversal, behind the reference's builder API.

Importing the package does not touch the GPU; the first Context() loads libbvh_cuda.so and fails loudly if the
library has not been built or no CUDA device is present (there is no CPU fallback)."""
from .types import BVH_NODE, TLAS_NODE, INSTANCE, MESH_INFO, MAX_DIST, NO_HIT  # noqa: F401
from ._lib import BvhCudaError, LIB_PATH  # noqa: F401
from .bvh import Bvh, BvhBuilder, Context, Ray, Scene, Tlas, Hit, MISS, default_context  # noqa: F401
from .bvh import gen_primary_rays_dev, gen_shadow_rays_dev, gen_area_shadow_rays_dev, instances_rotate_z_dev  # noqa: F401
