"""Data-parallel sharding of utterance batches: one process per GPU, no data-path collective.

The path has no cross-utterance reduction: utterances are dealt to ranks so
that the summed encoder length is balanced, each rank decodes its shard with its own handle, and
ONE collective at the end gathers fixed-size hypothesis records (tokens[max_len], len, score) -
`torch.distributed.all_gather` over NCCL (NVLink 5 / NVSwitch) on the GPUs, gloo in the CPU tests.
The reference has no distributed code at all."""
import numpy as np
import torch
import torch.distributed as dist


def shard_utterances(n_samples, world_size):
    """Balanced assignment: sort by length (descending) and deal round-robin, so every rank gets
    the same count (+-1) and a near-equal total duration.  Returns list[rank] -> sorted original
    indices."""
    order = np.argsort(-np.asarray(n_samples, dtype=np.int64), kind="stable")
    shards = [[] for _ in range(world_size)]
    for pos, idx in enumerate(order.tolist()):
        lap, slot = divmod(pos, world_size)
        rank = slot if lap % 2 == 0 else world_size - 1 - slot      # serpentine: balances totals
        shards[rank].append(idx)
    return [sorted(s) for s in shards]


def pack_records(indices, tokens, lens, scores, max_len):
    """[n, max_len + 3] int32 records: original index, len, score bits, tokens."""
    n = len(indices)
    rec = np.zeros((n, max_len + 3), dtype=np.int32)
    rec[:, 0] = np.asarray(indices, dtype=np.int32)
    rec[:, 1] = np.asarray(lens, dtype=np.int32)
    rec[:, 2] = np.asarray(scores, dtype=np.float32).view(np.int32)
    rec[:, 3:] = np.asarray(tokens, dtype=np.int32).reshape(n, max_len)
    return rec


def gather_hypotheses(rec, total, max_len, device=None, group=None):
    """All-gather the per-rank records (padded to the largest shard) and scatter them back into
    original utterance order.  Returns (tokens [total, max_len], lens [total], scores [total])."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = (total + world - 1) // world
    buf = np.full((per, max_len + 3), -1, dtype=np.int32)
    buf[:rec.shape[0]] = rec
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    if world > 1:
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        allrec = torch.cat(out, dim=0).cpu().numpy()
    else:
        allrec = t.cpu().numpy()
    allrec = allrec[allrec[:, 0] >= 0]
    tokens = np.zeros((total, max_len), dtype=np.int32)
    lens = np.zeros(total, dtype=np.int32)
    scores = np.zeros(total, dtype=np.float32)
    idx = allrec[:, 0]
    tokens[idx] = allrec[:, 3:]
    lens[idx] = allrec[:, 1]
    scores[idx] = allrec[:, 2].copy().view(np.float32)
    return tokens, lens, scores
