"""FWER-corrected (1-P) maps from the permutation maxima -- the array-level core of the reference's
tmanalysis/calculate_fweP_vertex.py:45-79 and calculate_fweP_voxel.py:33-63 (image file I/O through nibabel is
out of scope).

    corrp = fwe_corrected_p(perm_max, tfce_values)       # tfce_values > 0 only, like the reference's masks

`perm_max` is what np.genfromtxt reads from perm_*_TFCE_max{Vertex,Voxel}.csv.  Sorting the (few thousand)
maxima is host numpy exactly as in the reference; the per-vertex lookup runs on the GPU (tmb_fwe_lookup)."""
import numpy as np

from .. import _lib


def fwe_corrected_p(perm_max, tfce_values):
    import torch
    _lib.require_device()
    srt = np.sort(np.asarray(perm_max, dtype=np.float64).ravel())
    vals = np.ascontiguousarray(np.asarray(tfce_values, dtype=np.float32).ravel())
    dev = torch.device("cuda", torch.cuda.current_device())
    d_s = torch.from_numpy(srt).to(dev)
    d_v = torch.from_numpy(vals).to(dev)
    out = torch.empty((vals.shape[0],), dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().tmb_fwe_lookup(_lib.ptr(d_s), int(srt.shape[0]), _lib.ptr(d_v), int(vals.shape[0]),
                                         _lib.ptr(out), _lib.current_stream()))
    return out.cpu().numpy().reshape(np.shape(tfce_values))


def fwe_image(perm_max, tfce_image, neglog10=False):
    """Whole-image form: values <= 0 stay 0 (the reference only looks up `data > 0`)."""
    img = np.asarray(tfce_image, dtype=np.float32)
    out = np.zeros(img.shape, dtype=np.float64)
    m = img > 0
    out[m] = fwe_corrected_p(perm_max, img[m])
    if neglog10:
        with np.errstate(divide="ignore"):
            return -np.log10(1 - out)
    return out


def accuracy_line(num_perm):
    """calculate_fweP_vertex.py:75."""
    return "The accuracy is p = 0.05 +/- %.4f" % (2 * (np.sqrt(0.05 * 0.95 / num_perm)))
