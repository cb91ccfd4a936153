"""Host-side sharding of the hot path across the GPUs of one box: one process per GPU, no data-path collective.

* frames (tracked meshes = training images): rank r of R holds frames r, r + R, r + 2R, ... -- BVH, transform tables and images of the
  other frames never reach this GPU (ngp keeps all of them on one GPU, src/core.cu:294-334);
* ray / sample batches: every rank draws its own batch over ITS frames, so a global step is R x the per-rank batch ("weak" scaling);
* rendering: a frame renders on the rank that holds it;
* training: the only exchange is the sum of the gradient buffers; each rank scales its loss by 1 / R first, so the sum is the mean over
  the global batch and the fp16 sum cannot overflow where one rank's gradient would not (insta_train_nerf does this itself).
"""
import numpy as np


def frames_of_rank(n_frames, rank, world):
    """global frame ids held by `rank`"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_frames, world))


def owner_of_frame(frame, world):
    """(rank, local frame index) of a global frame id"""
    return frame % world, frame // world


def local_frame_ids(global_ids, rank, world):
    """per-sample `holder` values for this rank: local index of the samples whose frame lives here, -1 (= skipped by `warp`) otherwise"""
    g = np.asarray(global_ids, np.int64)
    return np.where((g >= 0) & (g % world == rank), g // world, -1).astype(np.int32)


def loss_scale_of_rank(loss_scale, world):
    """what every rank multiplies its per-ray loss gradients with, so that SUM over ranks == mean over the global batch"""
    return float(loss_scale) / float(world)


def render_schedule(frame_ids, world):
    """frames to render per rank, in the order given (each frame goes to its owner)"""
    out = [[] for _ in range(world)]
    for f in frame_ids:
        out[f % world].append(f)
    return out
