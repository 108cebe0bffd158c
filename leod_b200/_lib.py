"""ctypes binding of libleod_b200.so (the C ABI in include/leod_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LEOD_B200_LIB') or os.path.join(_HERE, 'lib', 'libleod_b200.so')      # the override serves kernel-variant experiments

LEOD_F32, LEOD_BF16, LEOD_U8 = 0, 1, 2
EPI_NONE, EPI_GELU, EPI_RESID, EPI_GELU_BWD = 0, 1, 2, 3


class BackboneCfg(Structure):
    _fields_ = [('in_channels', c_int32), ('embed_dim', c_int32), ('dim_head', c_int32), ('part_h', c_int32),
                ('part_w', c_int32), ('mlp_ratio', c_int32), ('in_h', c_int32), ('in_w', c_int32), ('dtype', c_int32),
                ('ln_eps', c_float)]


class DetectCfg(Structure):
    _fields_ = [('in_channels', c_int32 * 3), ('strides', c_int32 * 3), ('num_classes', c_int32), ('n_bottleneck', c_int32),
                ('in_h', c_int32), ('in_w', c_int32), ('dtype', c_int32), ('bn_eps', c_float), ('bn_momentum', c_float),
                ('ignore_label', c_float), ('n_ignore_thresh', c_int32), ('ignore_thresh', c_float * 8),
                ('reg_weight', c_float), ('obj_weight', c_float), ('cls_weight', c_float)]


class AugmState(Structure):
    """leod_augm_state (include/leod_b200.h): augmentation state of one sequence of the batch."""
    _fields_ = [('h_flip', c_int32), ('t_flip', c_int32), ('zoom_mode', c_int32), ('x0', c_int32), ('y0', c_int32),
                ('win_h', c_int32), ('win_w', c_int32), ('flip_c', c_float), ('lo_x', c_float), ('hi_x', c_float),
                ('lo_y', c_float), ('hi_y', c_float), ('mul', c_float), ('cap_x', c_float), ('cap_y', c_float)]


# int fn(void *ctx, double *buf, int64_t n, void *stream): sum buf over the ranks (leod_detect_set_allreduce)
ALLREDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_int64, c_void_p)

_VP4 = c_void_p * 4
_VP3 = c_void_p * 3
_lib = None


def _declare(lib):
    I, VP, F = c_int, c_void_p, c_float
    sig = {
        'leod_last_error': (c_char_p, []),
        'leod_abi_version': (I, []),
        'leod_launch_count': (ctypes.c_ulonglong, []),
        'leod_profile_enable': (I, [I]),
        'leod_debug_force_simt_attention': (I, [I]),
        'leod_profile_csv': (I, [c_char_p]),
        'leod_profile_collect': (I, [POINTER(ctypes.c_double), I]),
        'leod_backbone_create': (I, [POINTER(BackboneCfg), POINTER(VP)]),
        'leod_backbone_layout_only': (I, [POINTER(BackboneCfg), POINTER(VP)]),
        'leod_backbone_destroy': (None, [VP]),
        'leod_backbone_param_info': (I, [VP, I, c_char_p, c_size_t, POINTER(c_int64), POINTER(c_int32), POINTER(c_int64 * 4)]),
        'leod_backbone_param_count': (c_int64, [VP]),
        'leod_backbone_bind': (I, [VP, VP, VP]),
        'leod_backbone_prepare': (I, [VP, VP]),
        'leod_backbone_save_bytes': (c_int64, [VP, I]),
        'leod_backbone_reserve': (I, [VP, I]),
        'leod_backbone_set_gemm_impl': (I, [VP, I]),
        'leod_backbone_step_fwd': (I, [VP, VP, I, I, I, I, _VP4, _VP4, _VP4, _VP4, VP, VP]),
        'leod_backbone_step_bwd': (I, [VP, VP, I, I, I, I, _VP4, _VP4, _VP4, _VP4, VP, _VP4, _VP4, _VP4, _VP4, VP]),
        'leod_backbone_grads_finalize': (I, [VP, VP]),
        'leod_backbone_seq_arena_bytes': (c_int64, [VP, I, I]),
        'leod_backbone_seq_fwd': (I, [VP, VP, I, I, I, I, I, _VP4, _VP4, _VP4, _VP4, VP]),
        'leod_backbone_seq_bwd': (I, [VP, VP, I, I, I, I, I, _VP4, _VP4, _VP4, _VP4, _VP4, _VP4, _VP4, VP]),
        'leod_gemm_nt': (I, [I, I, VP, I, VP, I, I, VP, I, VP, I, I, I, I, VP, I, VP, I, VP, I, VP]),
        'leod_gemm_tn': (I, [I, I, VP, I, VP, I, VP, I, VP, I, I, I, VP]),
        'leod_stem_conv_fwd': (I, [VP, I, I, I, I, I, I, I, VP, I, VP, VP]),
        'leod_stem_conv_wgrad': (I, [VP, I, I, I, I, I, I, I, VP, VP, I, VP]),
        'leod_attention_fwd': (I, [I, VP, VP, I, I, I, I, I, I, I, I, VP]),
        'leod_attention_bwd': (I, [I, VP, VP, VP, I, I, I, I, I, I, I, I, VP]),
        'leod_layernorm_fwd': (I, [I, VP, VP, VP, VP, I, I, F, VP]),
        'leod_layernorm_bwd': (I, [I, VP, VP, VP, VP, VP, VP, VP, I, I, F, VP]),
        'leod_lstm_gates_fwd': (I, [I, VP, VP, VP, VP, I, I, VP]),
        'leod_lstm_gates_bwd': (I, [I, VP, VP, VP, VP, VP, VP, VP, VP, I, I, VP]),
        'leod_detect_create': (I, [POINTER(DetectCfg), POINTER(VP)]),
        'leod_detect_layout_only': (I, [POINTER(DetectCfg), POINTER(VP)]),
        'leod_detect_destroy': (None, [VP]),
        'leod_detect_param_info': (I, [VP, I, c_char_p, c_size_t, POINTER(c_int64), POINTER(c_int32), POINTER(c_int64 * 4)]),
        'leod_detect_buffer_info': (I, [VP, I, c_char_p, c_size_t, POINTER(c_int64), POINTER(c_int32), POINTER(c_int64 * 4)]),
        'leod_detect_counter_info': (I, [VP, I, c_char_p, c_size_t, POINTER(c_int64)]),
        'leod_detect_param_count': (c_int64, [VP]),
        'leod_detect_buffer_count': (c_int64, [VP]),
        'leod_detect_counter_count': (c_int64, [VP]),
        'leod_detect_num_anchors': (I, [VP]),
        'leod_detect_bind': (I, [VP, VP, VP, VP, VP]),
        'leod_detect_prepare': (I, [VP, VP]),
        'leod_detect_reserve': (I, [VP, I]),
        'leod_detect_set_allreduce': (I, [VP, ALLREDUCE_FN, VP]),
        'leod_fpn_head_fwd': (I, [VP, _VP3, I, I, VP, VP]),
        'leod_simota_loss_fwd': (I, [VP, VP, I, VP, VP]),
        'leod_simota_loss_bwd': (I, [VP, VP, I, VP, VP]),
        'leod_simota_assignment': (I, [VP, VP, VP, VP]),
        'leod_detect_get_raw': (I, [VP, VP, VP]),
        'leod_detect_get_raw_grad': (I, [VP, VP, VP]),
        'leod_detect_set_raw_grad': (I, [VP, VP, VP]),
        'leod_fpn_head_bwd': (I, [VP, _VP3, VP]),
        'leod_postprocess': (I, [VP, I, I, I, F, F, I, VP, VP, I, VP]),
        'leod_pred2label': (I, [VP, VP, I, I, I, POINTER(c_float), POINTER(c_float), I, I, VP, VP, VP]),
        'leod_tta_merge': (I, [VP, VP, I, I, F, F, I, VP, VP, VP]),
        'leod_track_workspace_bytes': (c_int64, [c_int64, I, I, I]),
        'leod_track_filter': (I, [VP, VP, VP, VP, VP, c_int64, I, I, POINTER(ctypes.c_double), I, ctypes.c_double, ctypes.c_double, F, I, I, I, F, I,
                                  VP, VP, VP, VP, VP, VP, VP, VP, VP]),
        'leod_pack_bbox': (I, [VP, c_int64, VP, I, VP]),
        'leod_upload_small': (I, [VP, VP, c_int64, VP]),
        'leod_augment_ev_repr': (I, [VP, VP, I, I, I, I, I, POINTER(AugmState), VP]),
        'leod_augment_labels': (I, [VP, VP, c_int64, I, POINTER(AugmState), VP, VP]),
        'leod_coco_eval_workspace_bytes': (c_int64, [I, I]),
        'leod_coco_eval': (I, [VP, VP, VP, VP, VP, VP, VP, VP, VP, I, I, c_int64, I, I, I, POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                               VP, VP, VP, VP, VP]),
        'leod_voxel_bin': (I, [VP, VP, VP, VP, c_int64, I, I, I, I, I, VP, VP]),
        'leod_adamw_ema': (I, [VP, VP, VP, VP, VP, c_int64, I, F, F, F, F, F, F, F, VP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED_SYMBOLS = ['leod_last_error', 'leod_abi_version', 'leod_launch_count', 'leod_profile_enable', 'leod_debug_force_simt_attention', 'leod_profile_csv',
                    'leod_profile_collect', 'leod_backbone_create', 'leod_backbone_layout_only',
                    'leod_backbone_destroy',
                    'leod_backbone_param_info', 'leod_backbone_param_count', 'leod_backbone_bind', 'leod_backbone_prepare',
                    'leod_backbone_save_bytes', 'leod_backbone_reserve', 'leod_backbone_set_gemm_impl',
                    'leod_backbone_step_fwd', 'leod_backbone_step_bwd', 'leod_backbone_grads_finalize', 'leod_backbone_seq_arena_bytes',
                    'leod_backbone_seq_fwd', 'leod_backbone_seq_bwd', 'leod_gemm_nt',
                    'leod_gemm_tn', 'leod_stem_conv_fwd', 'leod_stem_conv_wgrad', 'leod_attention_fwd', 'leod_attention_bwd', 'leod_layernorm_fwd', 'leod_layernorm_bwd',
                    'leod_lstm_gates_fwd', 'leod_lstm_gates_bwd', 'leod_detect_create', 'leod_detect_layout_only', 'leod_detect_destroy',
                    'leod_detect_param_info', 'leod_detect_buffer_info', 'leod_detect_counter_info', 'leod_detect_param_count',
                    'leod_detect_buffer_count', 'leod_detect_counter_count', 'leod_detect_num_anchors', 'leod_detect_bind',
                    'leod_detect_prepare', 'leod_detect_reserve', 'leod_detect_set_allreduce', 'leod_fpn_head_fwd',
                    'leod_simota_loss_fwd', 'leod_simota_loss_bwd', 'leod_simota_assignment', 'leod_detect_get_raw', 'leod_detect_get_raw_grad', 'leod_detect_set_raw_grad', 'leod_fpn_head_bwd', 'leod_postprocess', 'leod_pred2label', 'leod_tta_merge', 'leod_track_workspace_bytes', 'leod_track_filter', 'leod_pack_bbox',
                    'leod_upload_small', 'leod_augment_ev_repr', 'leod_augment_labels', 'leod_coco_eval_workspace_bytes', 'leod_coco_eval', 'leod_voxel_bin', 'leod_adamw_ema']


def lib():
    """Load (once) and return the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               f'or leod_b200/csrc/build.sh. There is no CPU/PyTorch fallback for the hot path.')
        l = ctypes.CDLL(LIB_PATH)
        _declare(l)
        _lib = l
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().leod_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'leod_b200 {what} failed ({rc}): {msg}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


class _Staging:
    """Ring of pinned host buffers for leod_upload_small.  A slot is reused only after the stream has passed the upload that read
    it (one event per slot; with 16 slots the wait never blocks in practice)."""
    SLOTS, SLOT_BYTES = 16, 1 << 18

    def __init__(self):
        self.buf = [torch.empty(self.SLOT_BYTES, dtype=torch.uint8).pin_memory() for _ in range(self.SLOTS)]
        self.ev = [None] * self.SLOTS
        self.keep = [None] * self.SLOTS
        self.next = 0

    def acquire(self):
        i = self.next
        self.next = (i + 1) % self.SLOTS
        if self.ev[i] is not None:
            self.ev[i].synchronize()
        return i


_staging = None


def upload_small(t: torch.Tensor, device) -> torch.Tensor:
    """Host tensor (a few KB: label rows, index lists, masks) -> new device tensor, enqueued on the current stream as an SM
    kernel (leod_upload_small), never on a copy engine.  The host is not blocked."""
    global _staging
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    padded = (nbytes + 3) // 4 * 4
    dst = torch.empty(max(padded, 4), dtype=torch.uint8, device=device)
    out = dst[:nbytes].view(t.dtype).view(t.shape)
    if nbytes == 0:
        return out
    if _staging is None:
        _staging = _Staging()
    slot = _staging.acquire()
    # larger than a slot: a private pinned buffer, kept alive by the slot's bookkeeping until the slot is reused
    host = _staging.buf[slot] if padded <= _Staging.SLOT_BYTES else torch.empty(padded, dtype=torch.uint8).pin_memory()
    host[:nbytes].copy_(t.view(-1).view(torch.uint8))
    with torch.cuda.device(dst.device):
        check(lib().leod_upload_small(ptr(dst), c_void_p(host.data_ptr()), padded, stream_ptr(dst.device)), 'upload_small')
        ev = torch.cuda.Event()
        ev.record()
    _staging.ev[slot] = ev
    _staging.keep[slot] = host
    return out


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def vp4(tensors):
    arr = _VP4()
    for i in range(4):
        t = tensors[i] if tensors is not None else None
        arr[i] = None if t is None else t.data_ptr()
    return arr


PROF_KINDS = ['gemm_nt', 'gemm_tn', 'attention_fwd', 'attention_bwd', 'layernorm', 'lstm_gates', 'patch', 'other', 'conv']


class _DevArray:
    """CUDA array interface over a raw device pointer (lets torch wrap library-owned memory without a copy)."""

    def __init__(self, ptr_, n, typestr):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (int(ptr_), False), 'version': 2}


def tensor_from_ptr(ptr_, n, dtype, device):
    typestr = {torch.float64: '<f8', torch.float32: '<f4', torch.int32: '<i4'}[dtype]
    return torch.as_tensor(_DevArray(ptr_, n, typestr), device=device)


def profile_collect():
    """-> {kind: dict(launches, ms, flops, bytes)} since the last call (needs leod_profile_enable(1))."""
    n = len(PROF_KINDS)
    buf = (ctypes.c_double * (4 * n))()
    check(lib().leod_profile_collect(buf, n), 'profile_collect')
    return {k: dict(launches=int(buf[4 * i]), ms=buf[4 * i + 1], flops=buf[4 * i + 2], bytes=buf[4 * i + 3])
            for i, k in enumerate(PROF_KINDS)}


def leod_dtype(dt):
    return {torch.float32: LEOD_F32, torch.bfloat16: LEOD_BF16, torch.uint8: LEOD_U8}[dt]
