"""STFT magnitude / phase front end: drop-in for utility_functions.spectrum_fast
(utility_functions.py:129-155), computed by the batched sm_100a kernel (csrc/stft.cuh)."""
import numpy as np
import torch

from . import functional as F


def spectrum_fast(x, nperseg=512, noverlap=128, window='hamming', cut_dc=True,
                  output_phase=True, cut_last_timeframe=True):
    """Same signature and output layout as the reference: (C, n) -> (C*(1+phase), F, T), magnitude
    planes first, then phase planes.  A leading batch dimension (B, C, n) is also accepted.
    numpy in -> numpy out (dtype of the input), CUDA tensor in -> CUDA tensor out (no host copy).
    Arithmetic is float32 on the GPU (the reference's float32 path differs from its float64
    path by 1.4e-7 relative)."""
    if window != 'hamming':
        raise NotImplementedError("seldq: only the Hamming window of the reference path is implemented")
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda()
        out = F.stft_magphase(t.float(), nperseg, noverlap, cut_dc, output_phase, cut_last_timeframe)
        return out if x.is_cuda else out.cpu().to(x.dtype)
    a = np.asarray(x)
    if a.ndim not in (2, 3):
        # the reference concatenates on axis -3, which only exists for (C, n) inputs
        raise ValueError("spectrum_fast expects (channels, samples)")
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    out = F.stft_magphase(t, nperseg, noverlap, cut_dc, output_phase, cut_last_timeframe)
    return out.cpu().numpy().astype(a.dtype if a.dtype.kind == 'f' else np.float32, copy=False)
