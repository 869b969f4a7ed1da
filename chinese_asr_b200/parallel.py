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
    """THE collective of the path: one all-gather of the per-rank records (padded to the
    largest shard) at the end of a run, scattered back into original utterance order.  Call it once with the
    records of every batch a rank decoded, not once per batch.  Returns (tokens [total, max_len], lens [total],
    scores [total])."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = (total + world - 1) // world
    buf = np.full((per, max_len + 3), -1, dtype=np.int32)
    buf[:rec.shape[0]] = rec
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.pin_memory().to(device, non_blocking=True)
    if world > 1:
        out = torch.empty((world * per, max_len + 3), dtype=torch.int32, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        allrec = out.cpu().numpy()
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


def pin_rank_to_local_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the CPU cores next to its GPU (NVML's ideal affinity of the
    device, split evenly between the ranks that share it), so that 8 ranks x (engine threads + NCCL proxy) do not
    migrate over all cores of the box.  Call before any worker thread exists; returns the cores or None when NVML
    or sched_setaffinity is unavailable or a rank's slice would be smaller than 4 cores (nothing is changed then).
    The caller must keep its OpenMP / torch thread count within the slice (bench.py: an os.cpu_count()-thread pool
    on a 1/8 slice made host-side torch work ~100x slower)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        words = (os.cpu_count() + 63) // 64
        sets = []
        for j in range(local_world):
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(j), words)
            sets.append(tuple(c for c in range(os.cpu_count()) if (mask[c // 64] >> (c % 64)) & 1))
        allowed = set(os.sched_getaffinity(0))
        mine = [c for c in sets[local_rank] if c in allowed] or sorted(allowed)
        sharing = [j for j in range(local_world) if sets[j] == sets[local_rank]]
        pos, n = sharing.index(local_rank), len(sharing)
        cores = mine[pos * len(mine) // n:(pos + 1) * len(mine) // n] or mine
        if len(cores) < 4:          # too few cores to be worth fencing in (main + 2 engine threads + NCCL proxy)
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


class BatchPipeline(object):
    """Several engines (handles) on ONE GPU, each driven by its own host thread and CUDA stream.

    A batch's encoder is latency-bound (4 x L dependent recurrence steps on 112 of the 148 SMs, at low
    utilisation); its decoder is throughput-bound.  With two batches in flight on two handles the encoder
    of one overlaps the decoder of the other: measured 28.4 -> 25.6 ms per batch of 512 x 10 s utterances
    (bw = 8) on one B200, hypotheses identical.  The C library needs nothing for this: handles are
    independent (one per thread, include/asr_b200.h), graph capture is thread-local.

    models: list of loaded `Model`s on the same device (each owns ~0.3 GB of weights + its workspace).
    submit() hands a batch to the next engine in round-robin order and returns a Future of what
    `Model.transcribe` returns; results keep submission order through the futures."""

    def __init__(self, models):
        from concurrent.futures import ThreadPoolExecutor
        if not models:
            raise ValueError("BatchPipeline needs at least one engine")
        self.models = list(models)
        self._pools = [ThreadPoolExecutor(max_workers=1) for _ in self.models]
        self._streams = [None] * len(self.models)
        self._next = 0

    def __len__(self):
        return len(self.models)

    def _run(self, i, fn):
        m = self.models[i]
        dev = getattr(m, "device", None)
        if dev is None or getattr(dev, "type", "cpu") != "cuda":
            return fn(m)                                    # host-only engines (tests)
        torch.cuda.set_device(dev)
        if self._streams[i] is None:
            self._streams[i] = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._streams[i]):
            return fn(m)

    def submit(self, pcm, offsets, **kw):
        i = self._next
        self._next = (self._next + 1) % len(self.models)
        return self._pools[i].submit(self._run, i, lambda m: m.transcribe(pcm, offsets, **kw))

    def map(self, batches, prefetch=False, **kw):
        """batches: list of (pcm, offsets) -> list of results in order.  prefetch=True (host buffers in pinned
        memory, all batches known in advance): every engine stages its NEXT batch's PCM on its copy stream
        (Model.prefetch / asr_prefetch_pcm) while it decodes the current one."""
        batches = list(batches)
        if kw.get("resident"):
            prefetch = False                                # the PCM is already in HBM: nothing to stage
        if not prefetch:
            futs = [self.submit(p, o, **kw) for p, o in batches]
            return [f.result() for f in futs]
        S = len(self.models)
        start = self._next
        self._next = (self._next + len(batches)) % S
        res = [None] * len(batches)

        def engine_loop(m, mine):
            bw = kw.get("bw")
            if mine:
                m.prefetch(batches[mine[0]][0], batches[mine[0]][1], bw=bw)
            for t, b in enumerate(mine):
                if t + 1 < len(mine):
                    m.prefetch(batches[mine[t + 1]][0], batches[mine[t + 1]][1], bw=bw)
                res[b] = m.transcribe(batches[b][0], batches[b][1], **kw)

        futs = []
        for j in range(S):
            i = (start + j) % S
            mine = list(range(j, len(batches), S))
            futs.append(self._pools[i].submit(self._run, i, lambda m, mine=mine: engine_loop(m, mine)))
        for f in futs:
            f.result()
        return res

    def each(self, fn):
        """Run fn(model) once on every engine's own thread / stream, one engine at a time (warm-up, reserve)."""
        return [self._pools[i].submit(self._run, i, fn).result() for i in range(len(self.models))]

    def close(self):
        for p in self._pools:
            p.shutdown(wait=True)
