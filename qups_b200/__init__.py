"""qups_b200 — B200-native (sm_100a) delay-and-sum beamformer and point-scatterer
simulator behind the API of thorstone25/qups (UltrasoundSystem.DAS / bfDAS /
greens -> kern/das_spec.m, kern/wsinterpd2.m, src/*.cu).

The compute lives in libqups_b200.so (hand-written CUDA, C ABI in
include/qups_b200.h); this package is the thin host-side mirror of the
reference's operator interface used by tests and bench.  No CPU fallback.
"""
from ._lib import QupsError, lib, launch_count, last_das_kernel, last_ws2_kernel, LIB_PATH  # noqa: F401
from .kern import das_spec, wsinterpd2, wsinterpd, convd, cohfac, dmas, pcf, slsc, pwznxcorr, FusedApod  # noqa: F401

__all__ = ["das_spec", "wsinterpd2", "wsinterpd", "convd", "cohfac", "dmas", "pcf", "slsc", "pwznxcorr", "FusedApod", "QupsError", "lib", "launch_count", "last_das_kernel", "last_ws2_kernel"]
