"""Multi-GPU partitioning of the LK path: independent frame pairs, contiguous blocks per rank, no data-path collective.

Every (frame pair) is independent and, once its pyramids exist, so is every feature.  Rank r of G owns the
pairs [floor(r*P/G), floor((r+1)*P/G)); each rank builds the pyramids of its own frames, so there is no duplicated work
and no inter-GPU traffic.  The only exchange is the final result gather (13 bytes per feature) to rank 0, which goes
through torch.distributed (NCCL on GPUs, gloo in the CPU tests) and is not on the critical path.
"""
import numpy as np


def shard_range(n_pairs, rank, world):
    """Contiguous block of pairs owned by `rank`: pair p lives on rank floor(p * world / n_pairs)."""
    lo = (rank * n_pairs) // world
    hi = ((rank + 1) * n_pairs) // world
    return lo, hi


def local_shard(prev, nxt, pts, offs, rank, world):
    """Slices a batch (prev/nxt (B,H,W), pts (N,2), offs (B+1,)) down to this rank's pairs; offsets are rebased to 0."""
    lo, hi = shard_range(len(offs) - 1, rank, world)
    p0, p1 = int(offs[lo]), int(offs[hi])
    return prev[lo:hi], nxt[lo:hi], pts[p0:p1], (np.asarray(offs[lo:hi + 1]) - p0).astype(np.int32), (lo, hi, p0, p1)


def track_sharded(track_fn, prev, nxt, pts, offs, rank, world, dist=None, device="cpu"):
    """Runs `track_fn(prev, nxt, pts, offs) -> (next_pts, status, err)` on this rank's block of pairs and gathers the
    results on rank 0 (returns (next_pts, status, err) there, None elsewhere).  `dist` is torch.distributed (already
    initialised) or None for a single process."""
    lp, ln, lpts, loffs, (lo, hi, p0, p1) = local_shard(prev, nxt, pts, offs, rank, world)
    if p1 > p0:
        npts, st, err = track_fn(lp, ln, lpts, loffs)
    else:
        npts, st, err = np.zeros((0, 2), np.float32), np.zeros(0, np.uint8), np.zeros(0, np.float32)
    if dist is None or world == 1:
        return npts, st, err
    import torch
    n_total = int(offs[-1])
    counts = [int(offs[shard_range(len(offs) - 1, r, world)[1]]) - int(offs[shard_range(len(offs) - 1, r, world)[0]]) for r in range(world)]
    cap = max(max(counts), 1)
    # fixed-size records (x, y, err, status) padded to the largest shard: one gather, 16 bytes per feature on the wire
    rec = torch.zeros((cap, 4), dtype=torch.float32, device=device)
    if p1 > p0:
        rec[:p1 - p0, 0:2] = torch.from_numpy(npts).to(device)
        rec[:p1 - p0, 2] = torch.from_numpy(err).to(device)
        rec[:p1 - p0, 3] = torch.from_numpy(st.astype(np.float32)).to(device)
    bufs = [torch.zeros_like(rec) for _ in range(world)] if rank == 0 else None
    dist.gather(rec, bufs, dst=0)
    if rank != 0:
        return None
    out_p, out_s, out_e = np.zeros((n_total, 2), np.float32), np.zeros(n_total, np.uint8), np.zeros(n_total, np.float32)
    pos = 0
    for r in range(world):
        b = bufs[r][:counts[r]].cpu().numpy()
        out_p[pos:pos + counts[r]] = b[:, 0:2]
        out_e[pos:pos + counts[r]] = b[:, 2]
        out_s[pos:pos + counts[r]] = b[:, 3].astype(np.uint8)
        pos += counts[r]
    return out_p, out_s, out_e
