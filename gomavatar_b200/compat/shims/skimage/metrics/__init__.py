"""``structural_similarity`` as eval.py:106-108 calls it — ``structural_similarity(pred, gt, multichannel=True)`` on
float64 ``[H,W,3]`` images that are 8-bit levels / 255 (eval.py:355-361) — evaluated by ``gom_eval_metrics``
(csrc/eval_metrics.cu: skimage 0.18 defaults, 7x7 uniform window, sample covariance, data_range 2 for float input,
mean over the cropped map and the channels; exact integer window sums, fp64 expression).  Other options are refused."""
import numpy as np


def structural_similarity(im1, im2, *, win_size=None, gradient=False, data_range=None, multichannel=False,
                          channel_axis=None, gaussian_weights=False, full=False, **kwargs):
    import torch
    from gomavatar_b200.metrics import eval_metrics
    if gradient or full or gaussian_weights or win_size not in (None, 7) or data_range not in (None, 2, 2.0) or kwargs:
        raise NotImplementedError("skimage stand-in: only the default options of eval.py:107 are implemented")
    if not (multichannel or channel_axis in (-1, 2)):
        raise NotImplementedError("skimage stand-in: [H,W,3] images with multichannel=True / channel_axis=-1 only")
    a, b = np.asarray(im1, dtype=np.float64), np.asarray(im2, dtype=np.float64)
    if a.shape != b.shape or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("structural_similarity: expected two [H,W,3] images")
    for x in (a, b):
        if np.abs(x * 255.0 - np.rint(x * 255.0)).max() > 1e-6 or x.min() < 0 or x.max() > 1:
            raise NotImplementedError("skimage stand-in: images must be 8-bit levels / 255 (to_8b_image(x) / 255.)")
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    m = eval_metrics(torch.from_numpy(a).float().to(dev)[None], torch.from_numpy(b).float().to(dev)[None], quantize=False)
    return float(m["ssim"][0])


def peak_signal_noise_ratio(image_true, image_test, *, data_range=None):
    a, b = np.asarray(image_true, dtype=np.float64), np.asarray(image_test, dtype=np.float64)
    if data_range is None:
        data_range = 2.0 if a.dtype.kind == "f" else 255.0
    return float(10 * np.log10(data_range ** 2 / np.mean((a - b) ** 2)))
