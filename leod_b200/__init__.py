"""leod_b200 — B200-native implementation of the LEOD hot path (RVT recurrent backbone, YOLOX head,
NMS, pseudo-label sweep) behind the reference's own model / module interface.

Sub-packages mirror the reference tree (`models/detection/...`, `modules/...`, `data/utils/...`) so a
reference checkout can switch by import path; the CUDA kernels live in `csrc/` behind the C ABI
declared in `include/leod_b200.h` and are loaded by `leod_b200._lib`.
"""
__version__ = '0.1.0'
