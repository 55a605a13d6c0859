"""Post-decode KITTI annotation conversion (host side, numpy) -- SURVEY.md §8(f) row 2, not yet built.

The reference does this on the CPU after the device->host copy (utils/kitti_convert_utils.py:16-249 with
utils/geometry_ops.py:7-93); it is outside the forward + decode hot path measured by bench.py.
"""


def convert_to_kitti_3d(results_3d, img_metas, calibs):
    raise NotImplementedError('KITTI annotation conversion is the next scope row (SURVEY.md §8f-2); '
                              'use batch_eval(..., get_vis_format=True) for the decoded boxes')


def convert_to_kitti_2d(results_2d, img_metas):
    raise NotImplementedError('KITTI annotation conversion is the next scope row (SURVEY.md §8f-2); '
                              'use batch_eval(..., get_vis_format=True) for the decoded boxes')
