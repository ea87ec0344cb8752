"""``apply_inverse_transforms`` (reference ``data_generator/object_detection_2d_misc_utils.py:22-73``): host-side list
plumbing around user-supplied inverter callables, same behaviour as the reference."""
import numpy as np


def apply_inverse_transforms(y_pred_decoded, inverse_transforms):
    if isinstance(y_pred_decoded, list):
        out = []
        for i in range(len(y_pred_decoded)):
            out.append(np.copy(y_pred_decoded[i]))
            if out[i].size > 0:
                for inverter in inverse_transforms[i]:
                    if inverter is not None:
                        out[i] = inverter(out[i])
        return out
    if isinstance(y_pred_decoded, np.ndarray):
        out = np.copy(y_pred_decoded)
        for i in range(len(y_pred_decoded)):
            if out[i].size > 0:
                for inverter in inverse_transforms[i]:
                    if inverter is not None:
                        out[i] = inverter(out[i])
        return out
    raise ValueError("`y_pred_decoded` must be either a list or a Numpy array.")
