"""Front end with train.py's surrounding steps fused in (csrc/stft.cuh): int16 PCM ingest, the
data-set normalisation of train.py:374-408 applied in the store, and the statistics pass that replaces np.mean /
np.std over the stored feature array.  Oracle: oracle/algebra.spectrum_fast (pinned to scipy / the reference in
tests/test_oracle.py) + numpy."""
import numpy as np
import pytest
import torch

from oracle import algebra as A

pytestmark = pytest.mark.gpu


def _clip(seed, B=2, C=8, n=32000):
    rng = np.random.default_rng(seed)
    return (0.1 * rng.standard_normal((B, C, n))).astype(np.float32)


@pytest.mark.parametrize("phase", [False, True])
def test_int16_ingest_matches_the_float_path(seldq, phase):
    x = _clip(1)
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    ref = np.stack([A.spectrum_fast(pcm[b].astype(np.float64) / 32768.0, nperseg=512, noverlap=112, output_phase=phase)
                    for b in range(x.shape[0])])
    out = seldq.functional.stft_features(torch.from_numpy(pcm).cuda(), 512, 112, True, phase, True)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert A.rel_err(o[:, :8], ref[:, :8]) < 1e-4
    if phase:
        # phase is ill-conditioned where the magnitude vanishes: compare through the unit phasors
        d = np.abs(np.exp(1j * o[:, 8:].astype(np.float64)) - np.exp(1j * ref[:, 8:]))
        assert float(np.quantile(d, 0.999)) < 1e-3


@pytest.mark.parametrize("phase", [False, True])
def test_statistics_pass_and_fused_normalisation_match_numpy(seldq, phase):
    """train.py:374-408: mean / std over the whole feature array per plane group, then (x - mean) / std.  Here: one
    statistics pass that stores nothing, then one pass that stores the normalised features."""
    x = _clip(2, B=3)
    feats = np.stack([A.spectrum_fast(x[b].astype(np.float64), nperseg=512, noverlap=112, output_phase=phase)
                      for b in range(x.shape[0])])
    groups = [feats[:, :8]] + ([feats[:, 8:]] if phase else [])
    xt = torch.from_numpy(x).cuda()
    stats = seldq.functional.stft_features(xt, 512, 112, True, phase, True, stats_only=True)
    ms = seldq.functional.feature_mean_std(stats[:len(groups)], groups[0].size)
    for (mu, sd), g in zip(ms, groups):
        assert abs(mu - g.mean()) < 1e-5 * max(1.0, abs(g.mean())) + 1e-7 and abs(sd - g.std()) < 1e-4 * g.std()
    out = seldq.functional.stft_features(xt, 512, 112, True, phase, True, mean_std=ms).cpu().numpy()
    ref = np.concatenate([(g - g.mean()) / g.std() for g in groups], axis=1)
    assert A.rel_err(out[:, :8], ref[:, :8]) < 1e-4
    assert abs(out[:, :8].mean()) < 1e-4 and abs(out[:, :8].std() - 1.0) < 1e-3


def test_device_submission_list_matches_the_reference_scan(seldq):
    """gen_submission_list_task2 (utility_functions.py:184-210) on the device against the restatement
    of the reference's triple loop (oracle/seld_metrics.events_per_frame, pinned to the imported reference in
    tests/test_oracle.py), including cells exactly at the 0.5 threshold (np.round: half to even -> inactive), frames
    without events, and the DCASE21 scores computed from both lists."""
    from oracle import seld_metrics as M
    rng = np.random.default_rng(3)
    clips, frames = 3, 600
    sed = rng.random((clips, frames, 42)).astype(np.float32) ** 6          # mostly inactive
    sed[:, ::7, 5] = 0.5                                                    # exactly at the threshold
    sed[:, 1::50, :] = 0.0                                                  # frames without events
    sed[0, 10, 3], sed[0, 10, 4] = 0.9, 0.51
    doa = (2 * rng.random((clips, frames, 126)) - 1).astype(np.float32)
    got = seldq.gen_submission_list_task2(torch.from_numpy(sed).cuda(), torch.from_numpy(doa).cuda())
    assert len(got) == clips
    for c in range(clips):
        want = M.events_per_frame(sed[c], doa[c])
        arr, d = got[c]
        assert sorted(d) == sorted(want)
        for fr in want:
            assert [e[0] for e in d[fr]] == [e[0] for e in want[fr]] and [e[4] for e in d[fr]] == [e[4] for e in want[fr]]
            assert np.allclose(np.array([e[1:4] for e in d[fr]]), np.array([e[1:4] for e in want[fr]]), rtol=1e-6, atol=1e-7)
        assert arr.shape == (sum(len(v) for v in want.values()), 5)
        assert np.all(np.diff(arr[:, 0]) >= 0)
    # single clip, numpy in: the reference's call signature
    arr1, d1 = seldq.gen_submission_list_task2(sed[1], doa[1], max_overlaps=3, max_loc_value=2.0)
    assert sorted(d1) == sorted(M.events_per_frame(sed[1], doa[1]))
