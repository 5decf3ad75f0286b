"""Host-side partitioning of the hot path across ranks (one process per GPU).

Two natural shardings, no invented collectives:

* ensembles -- independent trajectories (the reference's SLURM ``--array``, exampleSlurmFile.slurm:3, seeds
  ``time+job`` SU:1219): rank r advances jobs ``ensemble_jobs(n_jobs, world, r)``; nothing is exchanged.
* large N  -- i-row decomposition: rank r owns ions ``row_block(N, world, r)`` (their F rows, psi, V, tPart) and
  needs everybody's positions: ONE all-gather of ``3 * N / world`` doubles per rank per MD step, written in place
  into the ``[3][ld]`` position buffer (device memory of the engine on GPU ranks; any tensor in the gloo CPU tests).
"""
import numpy as np


def row_block(n_ions, world, rank):
    """(row0, n_rows) of rank `rank`: equal blocks; n_ions must divide evenly (all-gather of equal chunks)."""
    if n_ions % world:
        raise ValueError("n_ions=%d is not a multiple of world=%d (pad the system or pick another N)" % (n_ions, world))
    rows = n_ions // world
    return rank * rows, rows


def padded_ions(n_ions, world):
    """Smallest N' >= n_ions that row_block accepts."""
    return ((n_ions + world - 1) // world) * world


def ensemble_jobs(n_jobs, world, rank, first_job=1):
    """Job numbers (the reference's argv[1]) advanced by `rank`: contiguous, balanced to within one."""
    base, extra = divmod(n_jobs, world)
    start = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return list(range(first_job + start, first_job + start + count))


def allgather_positions(R, n_ions, world, rank, dist):
    """In-place all-gather of the row blocks of a ``[3][ld]`` position tensor (torch tensor, CPU or CUDA).

    Every rank contributes ``R[c, row0:row0+rows]`` and receives all N columns of each component. One collective per
    component (x, y, z); with NCCL the send buffer aliases its slot of the receive buffer (in-place all-gather)."""
    row0, rows = row_block(n_ions, world, rank)
    for c in range(3):
        out = R[c, :n_ions]
        if out.is_cuda:
            dist.all_gather_into_tensor(out, R[c, row0:row0 + rows])
        else:  # gloo: no aliasing of input and output
            dist.all_gather_into_tensor(out, R[c, row0:row0 + rows].clone())
    return R


def allreduce_scalars(values, dist, device=None):
    """Sum of per-rank partial observables (E_pot partial sums, kinetic sums, KDE bins) on output steps.
    NCCL needs device tensors: with that backend the values travel through the current CUDA device."""
    import torch
    t = torch.as_tensor(np.asarray(values, dtype=np.float64))
    if device is None and dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()


def distributed_diagnostics(engine, n_ions, dist, want_vel_dist=False):
    """output()'s observables (SU:934-979) of a row-decomposed run: two small all-reduces on output steps only.

    ``engine`` owns rows [row0,row0+n_rows) of one trajectory; returns the same dict on every rank."""
    s0 = allreduce_scalars(engine.diag_partial(None), dist)       # sum v_x (the energy slots of this call are unused)
    mean = s0[0] / n_ions
    s1 = allreduce_scalars(engine.diag_partial(mean), dist)
    out = {"vx_avg": mean, "ekin_x": s1[1] / n_ions, "ekin_y": s1[2] / n_ions, "ekin_z": s1[3] / n_ions, "epot": s1[4]}
    if want_vel_dist:
        out["pvel"] = allreduce_scalars(engine.vel_dist_partial(mean), dist)
    return out


# ---- the exchange of the row-decomposed step inside the library (csrc/mdqt_comm.cu), mirrored on the host for the CPU tests -------
def row_block_ceil(n_ions, world, rank):
    """(row0, n_rows, R) as mdqt_comm_init expects them: R = ceil(N / world) rows per rank, the last rank holds the remainder."""
    R = (n_ions + world - 1) // world
    row0 = rank * R
    return row0, max(0, min(R, n_ions - row0)), R


def pack_rows(X, row0, n_rows, R):
    """k_pack_rows: the rank's own rows of a ``[3][ld]`` array as one contiguous ``[3][R]`` block, zero padded."""
    blk = np.zeros((3, R), dtype=X.dtype)
    blk[:, :n_rows] = X[:, row0:row0 + n_rows]
    return blk


def unpack_rows(xbuf, X, n_ions, skip):
    """k_unpack_rows: ``xbuf[g][c][i]`` -> ``X[c][g*R + i]`` for every rank g but `skip`, rows beyond n_ions dropped."""
    G, _, R = xbuf.shape
    for g in range(G):
        if g == skip:
            continue
        lo, hi = g * R, min(n_ions, (g + 1) * R)
        if hi > lo:
            X[:, lo:hi] = xbuf[g, :, :hi - lo]
    return X


def exchange_rows(X, n_ions, world, rank, dist):
    """One in-place all-gather of a [3][R] block per rank into the rank-major buffer, then the unpack: what mdqt_md_steps does
    once per MD step with the fixed-point positions (ncclAllGather there, any torch.distributed backend here)."""
    import torch
    row0, n_rows, R = row_block_ceil(n_ions, world, rank)
    own = torch.from_numpy(np.ascontiguousarray(pack_rows(X, row0, n_rows, R))).reshape(-1)
    xbuf = torch.zeros(world * 3 * R, dtype=own.dtype)
    dist.all_gather_into_tensor(xbuf, own)
    return unpack_rows(xbuf.numpy().reshape(world, 3, R), X, n_ions, rank)
