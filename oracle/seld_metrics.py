"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Restatement of the reference's DCASE21 SELD metric pipeline, so that "SELD metrics identical on fixed seeds"
(BASELINE.json north_star) can be checked on the GPU box, where /root/reference does not exist:

    sed, doa --threshold at 0.5--> frame -> events      utility_functions.py:184-210 (gen_submission_list_task2)
             --1 s blocks-------> block -> class -> ... Dcase21_metrics.py:239-278  (segment_labels)
             --Hungarian matching, counting-----------> Dcase21_metrics.py:52-154   (SELDMetrics.update_seld_scores)
             --scores-----------> ER, F, LE, LR         Dcase21_metrics.py:32-50    (compute_seld_scores)

Coordinates are Cartesian (x, y, z), as gen_submission_list_task2 produces them; the polar branch of the reference
(:83-85, :212-214) is not restated.  The loop ORDER of the counting pass follows the reference (blocks, classes,
reference frames in order, tracks in first-match order), so the floating-point sums come out bit-identical; pinned
against the imported reference by tests/test_oracle.py on random multi-overlap inputs, and by the scores stored in
the model fixtures (oracle/make_golden.py computes those with the reference's own code).

Third-party arithmetic: scipy.optimize.linear_sum_assignment (Dcase21_metrics.py:218), numpy.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

_EPS = np.finfo(float).eps


def events_per_frame(sed, doa, max_loc_value=2.0, num_classes=14, max_overlaps=3):
    """utility_functions.py:184-210, the dictionary it returns: frame -> [[class, x, y, z, event number], ...] for
    every (class, overlap) cell whose SED output rounds to non-zero (np.round: half to even, i.e. > 0.5 ... and
    exactly 0.5 -> 0), in cell order; coordinates scaled back by max_loc_value."""
    sed = np.asarray(sed)
    doa = np.asarray(doa)
    out = {}
    for frame in range(sed.shape[0]):
        active = np.round(sed[frame])
        if np.sum(active) == 0:
            continue
        loc = (doa[frame] * max_loc_value).reshape(num_classes, max_overlaps, 3)
        for cell in np.flatnonzero(active != 0):
            cls, ev = int(cell // max_overlaps), int(cell % max_overlaps)
            out.setdefault(frame, []).append([cls, float(loc[cls][ev][0]), float(loc[cls][ev][1]), float(loc[cls][ev][2]), ev])
    return out


def blocks(frame_events, max_frames, frames_per_block=10):
    """Dcase21_metrics.py:239-278: block -> class -> [[frames within the block], [events of each of those frames]]
    (one such pair per class and block; an event is [x, y, z, event number])."""
    n_blocks = int(np.ceil(max_frames / float(frames_per_block)))
    out = {b: {} for b in range(n_blocks)}
    for start in range(0, max_frames, frames_per_block):
        per_class = {}
        for frame in range(start, start + frames_per_block):
            for ev in frame_events.get(frame, ()):
                per_class.setdefault(ev[0], {}).setdefault(frame - start, []).append(ev[1:])
        for cls, by_frame in per_class.items():
            out[start // frames_per_block].setdefault(cls, []).append([list(by_frame.keys()), list(by_frame.values())])
    return out


def _angular_distance_deg(a, b):
    """Dcase21_metrics.py:171-188: angle between Cartesian vectors, in degrees (normalisation with + 1e-10 under the root)."""
    na = np.sqrt(a[:, 0] ** 2 + a[:, 1] ** 2 + a[:, 2] ** 2 + 1e-10)
    nb = np.sqrt(b[:, 0] ** 2 + b[:, 1] ** 2 + b[:, 2] ** 2 + 1e-10)
    ax, ay, az, bx, by, bz = a[:, 0] / na, a[:, 1] / na, a[:, 2] / na, b[:, 0] / nb, b[:, 1] / nb, b[:, 2] / nb
    return np.arccos(np.clip(ax * bx + ay * by + az * bz, -1, 1)) * 180 / np.pi


def _match(gt_xyz, pred_xyz):
    """Dcase21_metrics.py:191-220: cost matrix of pairwise angular distances, Hungarian assignment."""
    g, p = gt_xyz.shape[0], pred_xyz.shape[0]
    cost = np.zeros((g, p))
    if g and p:
        gi, pi = np.meshgrid(np.arange(g), np.arange(p), indexing="xy")      # pairs in the reference's (pred-major) order
        gi, pi = gi.ravel(), pi.ravel()
        cost[gi, pi] = _angular_distance_deg(gt_xyz[gi], pred_xyz[pi])
    rows, cols = linear_sum_assignment(cost)
    return cost[rows, cols], rows, cols


class SeldScores(object):
    """Dcase21_metrics.py:4-154 (location-sensitive detection + class-sensitive localisation, 1 s segments)."""

    def __init__(self, doa_threshold=20, nb_classes=14):
        self.nb_classes = nb_classes
        self.threshold = doa_threshold
        self.TP = self.FP = self.FN = 0
        self.S = self.D = self.I = self.Nref = 0
        self.total_DE = 0
        self.DE_TP = self.DE_FP = self.DE_FN = 0

    def update(self, pred, gt):
        for b in range(len(gt.keys())):
            loc_FN = loc_FP = 0
            for cls in range(self.nb_classes):
                in_gt, in_pred = cls in gt[b], cls in pred[b]
                n_gt = max(len(v) for v in gt[b][cls][0][1]) if in_gt else None
                n_pred = max(len(v) for v in pred[b][cls][0][1]) if in_pred else None
                if in_gt:
                    self.Nref += n_gt
                if in_gt and in_pred:
                    dist_of_track, hits_of_track = {}, {}
                    gt_frames, pred_frames = gt[b][cls][0][0], pred[b][cls][0][0]
                    for k, frame in enumerate(gt_frames):
                        if frame not in pred_frames:
                            continue
                        g = np.array(gt[b][cls][0][1][k])
                        pk = pred_frames.index(frame)
                        q = np.array(pred[b][cls][0][1][pk])
                        dists, rows, _ = _match(g[:, :-1], q[:, :-1])
                        for d, track in zip(dists, rows):       # reference tracks are numbered by their position in the frame
                            dist_of_track.setdefault(track, []).append(d)
                            hits_of_track.setdefault(track, []).append(pk)
                    if not dist_of_track:
                        loc_FN += n_pred
                        self.FN += n_pred
                        self.DE_FN += n_pred
                    else:
                        for track in dist_of_track:
                            avg = sum(dist_of_track[track]) / len(hits_of_track[track])
                            self.total_DE += avg
                            self.DE_TP += 1
                            if avg <= self.threshold:
                                self.TP += 1
                            else:
                                loc_FP += 1
                                self.FP += 1
                        if n_pred > n_gt:
                            loc_FP += n_pred - n_gt
                            self.FP += n_pred - n_gt
                            self.DE_FP += n_pred - n_gt
                        elif n_pred < n_gt:
                            loc_FN += n_gt - n_pred
                            self.FN += n_gt - n_pred
                            self.DE_FN += n_gt - n_pred
                elif in_gt:
                    loc_FN += n_gt
                    self.FN += n_gt
                    self.DE_FN += n_gt
                elif in_pred:
                    loc_FP += n_pred
                    self.FP += n_pred
                    self.DE_FP += n_pred
            self.S += np.minimum(loc_FP, loc_FN)
            self.D += np.maximum(0, loc_FN - loc_FP)
            self.I += np.maximum(0, loc_FP - loc_FN)

    def scores(self):
        ER = (self.S + self.D + self.I) / float(self.Nref + _EPS)
        F = self.TP / (_EPS + self.TP + 0.5 * (self.FP + self.FN))
        LE = self.total_DE / float(self.DE_TP + _EPS) if self.DE_TP else 180
        LR = self.DE_TP / (_EPS + self.DE_TP + self.DE_FN)
        return ER, F, LE, LR


def seld_scores(sed, doa, target, n_sed=42, num_frames=None, num_classes=14, max_overlaps=3, doa_threshold=20,
                max_loc_value=2.0):
    """The evaluation loop of train.py:84-130 for a batch: sed (B, frames, n_sed), doa (B, frames, 3 n_sed), target
    (B, frames, 4 n_sed) with SED first (train.py:103-104) -> (ER, F, LE, LR) accumulated over the clips."""
    sed, doa, target = np.asarray(sed), np.asarray(doa), np.asarray(target)
    acc = SeldScores(doa_threshold, num_classes)
    for i in range(sed.shape[0]):
        frames = num_frames or sed.shape[1]
        pred = events_per_frame(sed[i], doa[i], max_loc_value, num_classes, max_overlaps)
        ref = events_per_frame(target[i][:, :n_sed], target[i][:, n_sed:], max_loc_value, num_classes, max_overlaps)
        acc.update(blocks(pred, frames), blocks(ref, frames))
    return tuple(float(v) for v in acc.scores())
