"""Post-processing of a generated batch, the steps tools/visualize.py:204-291 performs inline
between `model(**data)` and the SMPL-X export: body-part recomposition into the full pose vector and the
15 -> 30 fps up-sampling in 6D rotation space.  (The long-form cross-fade of longform_synthesis.py:449-471 is
`longform.crossfade_rotations`.)  Host-framework code: elementwise, runs on the tensors' device."""
import torch
import torch.nn.functional as F

from .longform import axis_angle_to_matrix, matrix_to_axis_angle, matrix_to_rotation_6d, rotation_6d_to_matrix


def recompose_motion(pred_upper, pred_lower, pred_hands, pred_face, upper_mask, lower_mask, hands_mask, face_mask):
    """[B, F, 39 / 27 / 90 / 3] axis-angle parts -> [B, F, len(mask)] full pose; the masks are the dataset's
    0/1 vectors over the 55*3 pose entries (beatx_dataset.py:82-109, tools/visualize.py:151-154,209-213)."""
    dev = pred_upper.device
    width = len(upper_mask)
    out = torch.zeros(pred_upper.shape[0], pred_upper.shape[1], width, dtype=pred_upper.dtype, device=dev)
    for part, mask in ((pred_upper, upper_mask), (pred_lower, lower_mask), (pred_hands, hands_mask), (pred_face, face_mask)):
        sel = torch.as_tensor(mask, device=dev).bool()
        if int(sel.sum()) != part.shape[-1]:
            raise ValueError(f"mask selects {int(sel.sum())} entries, the part has {part.shape[-1]}")
        out[..., sel] = part
    return out


def _interp(x, factor):
    return F.interpolate(x.permute(0, 2, 1), scale_factor=factor, mode="linear").permute(0, 2, 1)


def upsample_motion(motion_aa, facial, trans, motion_fps, target_fps=30):
    """Rotations are interpolated linearly in 6D space and projected back (visualize.py:262-291); expressions
    and translation linearly.  [B, F, J*3], [B, F, 100], [B, F, 3] -> the same with F * target_fps/motion_fps."""
    if target_fps == motion_fps:
        return motion_aa, facial, trans
    if target_fps % motion_fps:
        raise ValueError("target_fps must be a multiple of motion_fps")
    k = target_fps / motion_fps
    B, n, dim = motion_aa.shape
    J = dim // 3
    six = matrix_to_rotation_6d(axis_angle_to_matrix(motion_aa.reshape(B, n, J, 3))).reshape(B, n, J * 6)
    six = _interp(six, k)
    aa = matrix_to_axis_angle(rotation_6d_to_matrix(six.reshape(B, six.shape[1], J, 6))).reshape(B, six.shape[1], J * 3)
    return aa, _interp(facial, k), _interp(trans, k)
