"""Model configuration helpers.

The model classes take the reference's config nodes (OmegaConf DictConfig: attribute access + .get,
see config/model/maxvit_yolox/default.yaml:1-64).  `make_model_cfg` builds an equivalent node without
omegaconf/hydra for the benchmark, the tests and standalone use; values follow
config/experiment/gen{1,4}/{tiny,small,base}.yaml and config/modifier.py:49-64.
"""


class Node(dict):
    """Attribute-access dict with .get — the subset of DictConfig the model code uses."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = Node(v) if isinstance(v, dict) and not isinstance(v, Node) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


SIZES = {'tiny': (32, 32, 0.33), 'small': (48, 24, 0.33), 'base': (64, 32, 0.67)}
DATASETS = {'gen1': dict(partition_size=(8, 10), num_classes=2, in_res_hw=(256, 320), frame_hw=(240, 304)),
            'gen4': dict(partition_size=(6, 10), num_classes=3, in_res_hw=(384, 640), frame_hw=(360, 640))}


def _default_thresh(num_classes):
    """gen1 (car, ped) -> [0.6, 0.3]; gen4 (ped, cyc, car) -> [0.3, 0.3, 0.6] would need the dataset's class order:
    the reference re-orders at run time (config/modifier.py:88-108); here extra classes just take 0.3."""
    return ([0.6, 0.3] + [0.3] * max(0, num_classes - 2))[:num_classes]


def make_model_cfg(size=None, dataset=None, embed_dim=None, dim_head=None, partition_size=None, num_classes=None,
                   fpn_depth=None, input_channels=20, in_res_hw=None, ignore_bbox_thresh=None, compute_dtype='bf16',
                   conf_thre=0.1, nms_thre=0.45, pseudo_label=None):
    if size is not None:
        e, dh, dep = SIZES[size]
        embed_dim = embed_dim or e
        dim_head = dim_head or dh
        fpn_depth = fpn_depth or dep
    if dataset is not None:
        ds = DATASETS[dataset]
        partition_size = partition_size or ds['partition_size']
        num_classes = num_classes or ds['num_classes']
        in_res_hw = in_res_hw or ds['in_res_hw']
    if in_res_hw is None:
        in_res_hw = (32 * partition_size[0], 32 * partition_size[1])
    return Node(
        backbone=dict(name='MaxViTRNN', compile=dict(enable=False, args=dict(mode='reduce-overhead')),
                      input_channels=input_channels, enable_masking=False, partition_split_32=1, embed_dim=embed_dim,
                      dim_multiplier=(1, 2, 4, 8), num_blocks=(1, 1, 1, 1), T_max_chrono_init=(4, 8, 16, 32),
                      stem=dict(patch_size=4), in_res_hw=tuple(in_res_hw), compute_dtype=compute_dtype,
                      stage=dict(downsample=dict(type='patch', overlap=True, norm_affine=True),
                                 attention=dict(use_torch_mha=False, partition_size=tuple(partition_size), dim_head=dim_head,
                                                attention_bias=True, mlp_activation='gelu', mlp_gated=False, mlp_bias=True,
                                                mlp_ratio=4, drop_mlp=0, drop_path=0, ls_init_value=1e-5),
                                 lstm=dict(dws_conv=False, dws_conv_only_hidden=True, dws_conv_kernel_size=3,
                                           drop_cell_update=0))),
        fpn=dict(name='PAFPN', compile=dict(enable=False, args={}), depth=fpn_depth, in_stages=(2, 3, 4), depthwise=False,
                 act='silu'),
        head=dict(name='YoloX', compile=dict(enable=False, args={}), depthwise=False, act='silu', num_classes=num_classes,
                  obj_focal_loss=False, bbox_loss_weighting='', ignore_bbox_thresh=ignore_bbox_thresh, ignore_label=1024,
                  ignore_bg_k=0),
        postprocess=dict(confidence_threshold=conf_thre, nms_threshold=nms_thre),
        # config/model/pseudo_labeler.yaml:15-21 (read by modules/pseudo_labeler.py only)
        pseudo_label=dict(pseudo_label or dict(skip_first_t=0, obj_thresh=_default_thresh(num_classes),
                                               cls_thresh=_default_thresh(num_classes))))
