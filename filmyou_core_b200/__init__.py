"""filmyou_core_b200 -- B200-native engine for filmyou-core's RM2 hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/filmyou_rm2.h),
`engine` (ctypes binding of libfilmyou_rm2.so), `rm2_job` (host-side mirror of the reference's
RM2Job / reducer sink interface), `nmf` (the PPC / NMF clustering step that produces `clustering`), `baseline_job` (the options of config 3's similarity phase around the co-occurrence count), `sharding` (one process per GPU) and `datagen` (synthetic inputs).
There is no CPU fallback: importing works anywhere, computing needs a B200 and the built library.
"""
from .engine import Rm2Engine, Rm2Params, Rm2Error, library_path, load_library, build_library  # noqa: F401
from .nmf import NmfEngine, cluster_users  # noqa: F401,E402
