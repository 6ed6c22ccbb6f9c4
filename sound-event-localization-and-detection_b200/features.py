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


def gen_submission_list_task2(sed, doa, max_loc_value=2., num_frames=600, num_classes=14, max_overlaps=3):
    """Drop-in for utility_functions.gen_submission_list_task2 (utility_functions.py:184-210): the list of active sounds
    [frame, class, x, y, z] and the frame -> [[class, x, y, z, event number], ...] dictionary, from the model's SED /
    DOA outputs.  The frame x class x overlap scan runs in one CUDA kernel (csrc/eval.cu); only the active rows come
    back to the host.  sed (frames, classes * overlaps) and doa (frames, 3 * classes * overlaps) as in the reference, or
    with a leading clip dimension -- then a list with one (array, dict) pair per clip is returned.  numpy arrays or
    tensors; like the reference, `num_frames` is not used (the inputs' own frame count is)."""
    import ctypes
    from . import _lib
    ts, td = (t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)) for t in (sed, doa))
    batched = ts.dim() == 3
    ts, td = (t if batched else t[None] for t in (ts, td))
    ts, td = ts.detach().cuda().float().contiguous(), td.detach().cuda().float().contiguous()
    clips, frames, cells = ts.shape
    if cells != num_classes * max_overlaps or td.shape != (clips, frames, 3 * cells):
        raise ValueError("sed must be (frames, %d) and doa (frames, %d)" % (num_classes * max_overlaps, 3 * num_classes * max_overlaps))
    rows = torch.empty((clips, frames * cells, 6), dtype=torch.float32, device=ts.device)
    counts = torch.empty((clips,), dtype=torch.int32, device=ts.device)
    with torch.cuda.device(ts.device):
        _lib.check(_lib.lib().seldq_seld_events(ts.data_ptr(), td.data_ptr(), clips, frames, num_classes, max_overlaps,
                                                float(max_loc_value), rows.data_ptr(), counts.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    counts = counts.cpu().tolist()
    out = []
    for c in range(clips):
        r = rows[c, :counts[c]].cpu().numpy().astype(np.float64)
        d = {}
        for fr, cls, x, y, z, ev in r.tolist():
            d.setdefault(int(fr), []).append([int(cls), x, y, z, int(ev)])
        out.append((r[:, :5].copy() if counts[c] else np.array([]), d))
    return out if batched else out[0]
