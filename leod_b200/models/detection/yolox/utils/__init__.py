from .boxes import bboxes_iou, postprocess  # noqa: F401
