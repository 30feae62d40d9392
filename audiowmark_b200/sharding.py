"""Multi-GPU sharding of one long stream (one process per GPU, torch.distributed).

`get` shards by the reference's own chunks (WavChunkLoader: 30 min, 134.4 s overlap; src/wavchunkloader.cc:54-163) so
every per-chunk statistic (local mean, n-best, "all" pattern) is identical to the single-process run; `add` shards by
frame blocks with a recomputed halo (one 1024-frame for the synthesis window tails, two limiter blocks for the gain ramp)
so the interior of every shard is identical to the unsharded output.  Nothing is exchanged on the data path; the only
collective is the final gather of the chunk results (a few hundred bytes per detection).
"""
from __future__ import annotations

import os
import struct

FRAME = 1024


def chunk_plan(n_frames: int, max_frames: int, overlap: int, rate: int = 44100):
    """[(first_frame, n_frames_in_chunk, time_offset_seconds)] exactly as get_watermark_buffer walks a stream."""
    if n_frames <= 0:
        return []
    out = []
    start, end, toff = 0, min(max_frames, n_frames), 0.0
    eof = end < max_frames
    while True:
        out.append((start, end - start, toff))
        if eof:
            break
        toff += float(end - start - overlap) / rate
        start = end - overlap
        new_end = min(start + max_frames, n_frames)
        eof = (new_end - start) < max_frames
        end = new_end
    return out


def assign_chunks(n_chunks: int, world: int):
    """contiguous, balanced: rank r gets chunks [lo, hi)."""
    base, extra = divmod(n_chunks, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def embed_range(own_start: int, own_end: int, n_frames: int, limiter_block: int):
    """Extended range [ext_start, ext_end) a rank has to embed so that [own_start, own_end) comes out exactly as in the
    unsharded run.  ext_start is a multiple of 1024 (-> first_frame_number); returns (ext_start, ext_end, first_frame_number)."""
    if own_start <= 0:
        ext_start = 0
    else:
        b = own_start // limiter_block                     # first owned limiter block
        lo = max(b - 1, 0) * limiter_block                 # the block before it must be complete
        ext_start = max((lo // FRAME) * FRAME - FRAME, 0)  # + one frame for the synthesis-window tail
    if own_end >= n_frames:
        ext_end = n_frames
    else:
        b = (own_end - 1) // limiter_block                 # last owned limiter block
        hi = (b + 2) * limiter_block                       # the block after it must be complete
        ext_end = ext_start + -(-(hi - ext_start) // FRAME) * FRAME + FRAME
        ext_end = min(ext_end, n_frames)
    return ext_start, ext_end, ext_start // FRAME


def rank_ranges(n_frames: int, rank: int, world: int, max_frames: int, overlap: int, rate: int = 44100):
    """chunks of this rank and the PCM range [lo, hi) they cover."""
    plan = chunk_plan(n_frames, max_frames, overlap, rate)
    lo_c, hi_c = assign_chunks(len(plan), world)[rank]
    mine = plan[lo_c:hi_c]
    if not mine:
        return plan, (lo_c, hi_c), None
    return plan, (lo_c, hi_c), (mine[0][0], mine[-1][0] + mine[-1][1])


# ---- gathering chunk results ---------------------------------------------------------------------------------------

def pack_blobs(blobs) -> bytes:
    """[(chunk_index, bytes)] -> one byte string."""
    out = struct.pack("<I", len(blobs))
    for idx, b in blobs:
        out += struct.pack("<II", idx, len(b)) + b
    return out


def unpack_blobs(data: bytes):
    n, = struct.unpack_from("<I", data, 0)
    pos, out = 4, []
    for _ in range(n):
        idx, ln = struct.unpack_from("<II", data, pos)
        pos += 8
        out.append((idx, bytes(data[pos:pos + ln])))
        pos += ln
    return out


def gather_blobs(blobs, device=None):
    """all ranks -> list over ranks of [(chunk_index, bytes)].  Uses tensor collectives (all_gather of sizes, then of
    padded byte tensors) so it runs on NCCL (device = cuda) and on gloo (device = None / cpu)."""
    import torch
    import torch.distributed as dist
    payload = pack_blobs(blobs)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [unpack_blobs(payload)]
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    size = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    cap = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    return [unpack_blobs(bytes(b[:int(s.item())].cpu().numpy().tobytes())) for b, s in zip(bufs, sizes)]


# =====================================================================================================================
# Frame-balanced `get` across ranks.
#
# Chunk sharding above is exact but coarse (30-minute units).  Here every rank owns an equal span of stream POSITIONS;
# for each chunk that overlaps its span it runs the GPU stages on the slice of the chunk's start frames it owns (with a
# 6-frame margin for the local mean and a 2227-frame tail the sync pattern reaches into), and the per-chunk decisions
# (candidate selection, threshold / n-best, AB / "all" combination, merge) are taken on the gathered lists by every rank
# identically.  Results are identical to the single-process run; four small gathers (peaks, refined scores, soft bits,
# decoded words) are the only communication.
# =====================================================================================================================

import ctypes

import numpy as np

T_BLOCK = 2226                 # frames per block (payload 128 bit): sync pattern length in BLOCK mode
MARGIN = 6                     # start frames of margin on each side of a slice (local mean uses +-20 scores = +-5 frames)


def owner_span(n_total: int, world: int) -> int:
    per = -(-n_total // world)
    return -(-per // FRAME) * FRAME


class Slice:
    __slots__ = ("chunk", "sa", "sb", "a", "b", "lo", "hi")

    def __init__(self, chunk, sa, sb, a, b, lo, hi):
        self.chunk, self.sa, self.sb, self.a, self.b, self.lo, self.hi = chunk, sa, sb, a, b, lo, hi


def rank_slices(plan, rank: int, world: int, n_total: int):
    """slices (chunk, owned start frames [sa,sb), searched [a,b), stream PCM range [lo,hi)) of one rank"""
    span = owner_span(n_total, world)
    out = []
    for c, (cs, cn, _) in enumerate(plan):
        n_starts = max(cn // FRAME - T_BLOCK - 1, 0)
        if n_starts == 0:
            continue
        # start frame s sits at stream position cs + s*1024; owner = position // span (last rank takes the remainder)
        def first_s(pos):                       # smallest s with cs + s*1024 >= pos
            return max(0, -(-(pos - cs) // FRAME))
        sa = min(first_s(rank * span), n_starts)
        sb = n_starts if rank == world - 1 else min(first_s((rank + 1) * span), n_starts)
        if sb <= sa:
            continue
        a, b = max(sa - MARGIN, 0), min(sb + MARGIN, n_starts)
        hi = cs + cn if b == n_starts else cs + (b + T_BLOCK + 1) * FRAME     # the last slice keeps the chunk's partial tail frame
        out.append(Slice(c, sa, sb, a, b, cs + a * FRAME, hi))
    return out


def index_owners(plan, n_total: int, world: int, chunk: int, indices) -> np.ndarray:
    """rank that owns each chunk-relative sample index (by the start frame it falls into): the rank whose slice of `chunk`
    contains that start frame in [sa, sb) (rank_slices uses the same position rule)"""
    cs, cn, _ = plan[chunk]
    n_starts = max(cn // FRAME - T_BLOCK - 1, 0)
    s = np.clip(np.asarray(indices, np.int64) // FRAME, 0, max(n_starts - 1, 0))
    span = owner_span(n_total, world)
    return np.minimum((cs + s * FRAME) // span, world - 1)


def _pack(arrs) -> bytes:
    import pickle
    return pickle.dumps(arrs, protocol=4)


def _unpack(b: bytes):
    import pickle
    return pickle.loads(b)


_AG = {}


def allgather_bytes(payload: bytes, device=None, cap: int = 1 << 19):
    """bytes from every rank with ONE tensor collective (works on NCCL and gloo): every rank contributes a fixed-size
    slot [u64 length | payload | padding]; persistent device / pinned buffers.  Payloads larger than the slot fall back
    to a second, exactly sized exchange."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [payload]
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    key = (str(dev), world, cap)
    if key not in _AG:
        pin = dev.type == "cuda"
        _AG[key] = (torch.zeros(cap, dtype=torch.uint8, device=dev), torch.zeros(world * cap, dtype=torch.uint8, device=dev),
                    torch.zeros(cap, dtype=torch.uint8, pin_memory=pin), torch.zeros(world * cap, dtype=torch.uint8, pin_memory=pin))
    send, recv, hsend, hrecv = _AG[key]
    n = len(payload)
    fits = n + 8 <= cap
    hsend[:8] = torch.frombuffer(bytearray(struct.pack("<Q", n)), dtype=torch.uint8)
    if fits and n:
        hsend[8:8 + n] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    send.copy_(hsend, non_blocking=True)
    dist.all_gather_into_tensor(recv, send)
    hrecv.copy_(recv)
    raw = hrecv.numpy()
    sizes = [struct.unpack_from("<Q", raw, r * cap)[0] for r in range(world)]
    if all(sz + 8 <= cap for sz in sizes):
        return [raw[r * cap + 8:r * cap + 8 + sizes[r]].tobytes() for r in range(world)]
    big = max(sizes)                                    # rare: someone had more than a slot's worth
    buf = torch.zeros(big, dtype=torch.uint8, device=dev)
    if n:
        buf[:n] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    return [bytes(b[:sz].cpu().numpy().tobytes()) for b, sz in zip(bufs, sizes)]


class BalancedGet:
    """One rank of the frame-balanced `get`.  `pcm` is this rank's part of the (marked) stream: a numpy array [n, ch] or a
    device pointer, starting at stream frame `pcm_start`.  Stages alternate with gathers (see run())."""

    def __init__(self, rank, world, n_total, pcm, pcm_start, pcm_frames, channels, key=None, sample_rate=44100):
        from . import capi, hostapi as H
        self.H, self.capi = H, capi
        self.rank, self.world, self.n_total = rank, world, n_total
        self.pcm, self.pcm_start, self.pcm_frames, self.ch, self.rate = pcm, pcm_start, pcm_frames, channels, sample_rate
        if isinstance(pcm, np.ndarray):
            # host audio (float32 or 16 bit PCM): one upload for all stages -- every stage binds slices of the same device copy
            import torch
            d = torch.from_numpy(np.ascontiguousarray(pcm)).to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
            if d.dtype == torch.int16:
                d = d.to(torch.float32) * (1.0 / 32768.0)     # exact; the reference's int -> float rule (src/sfinputstream.cc:189-210)
            elif d.dtype != torch.float32:
                d = d.to(torch.float32)
            torch.cuda.current_stream().synchronize()
            self._device_copy = d
            self.pcm = d.data_ptr()
        self.key = bytes(key) if key is not None else bytes(16)
        mx, ov = H.chunk_geometry(sample_rate)
        self.plan = chunk_plan(n_total, mx, ov, sample_rate)
        self.slices = rank_slices(self.plan, rank, world, n_total)
        self.ctx = capi.Context.from_handle(H.engine_ctx())
        self.slot = H.key_slot(self.key)
        self.thr1 = H.get_param("sync_threshold2") * 0.75
        self.n_coded = H.n_coded_bits()

    # -- helpers
    def _bind(self, sl: Slice):
        off = sl.lo - self.pcm_start
        n = sl.hi - sl.lo
        assert off >= 0 and off + n <= self.pcm_frames, (off, n, self.pcm_frames)
        if isinstance(self.pcm, np.ndarray):
            self.ctx.pcm_bind(self.pcm[off:off + n])
        else:
            self.ctx.pcm_bind(int(self.pcm) + off * self.ch * 4, n, self.ch)

    def _slice_of(self, chunk, start_frame):
        for sl in self.slices:
            if sl.chunk == chunk and sl.a <= start_frame < max(sl.b, sl.sb + 1):
                if sl.sa <= start_frame < sl.sb or (start_frame >= sl.sb and sl.b == sl.sb) or (start_frame < sl.sa and sl.a == sl.sa):
                    return sl
        return None

    def owner_of(self, chunk, index):
        """rank that owns chunk-relative sample index `index` (by the start frame it falls into)"""
        cs, cn, _ = self.plan[chunk]
        n_starts = max(cn // FRAME - T_BLOCK - 1, 0)
        s = min(max(index // FRAME, 0), max(n_starts - 1, 0))
        span = owner_span(self.n_total, self.world)
        return min((cs + s * FRAME) // span, self.world - 1)

    def _my_slice(self, chunk):
        for sl in self.slices:
            if sl.chunk == chunk:
                return sl
        return None

    # -- stage 1: approximate search on my slices -> peaks above an (adaptive) floor, in chunk coordinates
    def stage_peaks(self, floors=None) -> bytes:
        out = []
        for sl in self.slices:
            self._bind(sl)
            self.ctx.sync_approx_run(self.slot, self.capi.MODE_BLOCK)
            seq = [self.thr1, self.thr1 * 0.6, self.thr1 * 0.35, self.thr1 * 0.15, -1.0]
            if floors and sl.chunk in floors:
                seq = [floors[sl.chunk]]
            for f in seq:
                pk, n = self.ctx.sync_peaks(f, 1 << 15 if f >= 0 else 1 << 17)
                if n > len(pk):
                    raise RuntimeError("too many peaks above floor %g" % f)
                own = pk[(pk["index"] >= (sl.sa - sl.a) * FRAME) & (pk["index"] < (sl.sb - sl.a) * FRAME)].copy()
                if len(own) >= 64 or f < 0:
                    break
            own["index"] += sl.a * FRAME
            out.append((sl.chunk, f, own))
        return _pack(out)

    # -- stage 2: candidate selection per chunk from everybody's peaks (identical on every rank)
    def stage_select(self, payloads):
        per_chunk = {}
        for p in payloads:
            for chunk, floor_q, pk in _unpack(p):
                per_chunk.setdefault(chunk, []).append((floor_q, pk))
        self.cands, retry = {}, {}
        for chunk, lst in per_chunk.items():
            floor_q = max(f for f, _ in lst)
            pk = np.concatenate([a for _, a in lst]) if lst else np.zeros(0, self.capi.SEARCH_SCORE)
            pk = pk[np.argsort(pk["index"], kind="stable")]
            sel, complete = self.H.stage_select(pk, floor_q)
            if not complete and floor_q >= 0:
                retry[chunk] = -1.0
            self.cands[chunk] = sel
        return retry            # chunks whose peak lists were too short (rare): ask for all peaks and select again

    def owners(self, chunk, indices) -> np.ndarray:
        """owner_of for an array of chunk-relative sample indices"""
        return index_owners(self.plan, self.n_total, self.world, chunk, indices)

    def viterbi_rank(self, chunk) -> int:
        """all code words of a chunk are decoded (and packed for the merge) by one rank, so that no rank builds the job list of
        every chunk of the stream"""
        return chunk % self.world

    # -- stage 3: refine the candidates I own
    def stage_refine(self) -> bytes:
        out = []
        for chunk, sel in self.cands.items():
            mine = np.nonzero(self.owners(chunk, sel["index"]) == self.rank)[0].tolist() if len(sel) else []
            sl = self._my_slice(chunk)
            if not mine or sl is None:
                continue
            self._bind(sl)
            part = sel[mine].copy()
            part["index"] -= sl.a * FRAME
            ref = self.ctx.sync_refine(part, self.slot, self.capi.MODE_BLOCK)
            ref["index"] += sl.a * FRAME
            out.append((chunk, np.array(mine, np.int64), ref))
        return _pack(out)

    # -- stage 4: threshold2 / n-best per chunk
    def stage_final(self, payloads):
        refined = {c: sel.copy() for c, sel in self.cands.items()}
        for p in payloads:
            for chunk, pos, ref in _unpack(p):
                refined[chunk][pos] = ref
        self.final = {c: self.H.stage_final(r) for c, r in refined.items()}       # (index, quality, btype) arrays

    # -- stage 5: soft bits of the final scores I own
    def stage_decode(self) -> bytes:
        out = []
        for chunk, (idx, q, bt) in self.final.items():
            mine = np.nonzero(self.owners(chunk, idx) == self.rank)[0].tolist() if len(idx) else []
            sl = self._my_slice(chunk)
            if not mine or sl is None:
                continue
            self._bind(sl)
            rel = idx[mine].astype(np.int64) - sl.a * FRAME
            # fft_range validity is decided against the CHUNK length: the slice either reaches the chunk end or is long enough
            raw, valid = self.ctx.decode_blocks(np.maximum(rel, 0).astype(np.uint64), self.n_coded, self.slot)
            out.append((chunk, np.array(mine, np.int64), raw, valid))
        return _pack(out)

    # -- stage 6: Viterbi: the code words of chunk c are built, decoded and packed by rank viterbi_rank (c)
    def stage_viterbi(self, payloads) -> bytes:
        import struct as st
        my_chunks = [c for c in sorted(self.final) if self.viterbi_rank(c) == self.rank]
        raws = {c: (np.zeros((len(self.final[c][0]), self.n_coded), np.float32), np.zeros(len(self.final[c][0]), np.int32)) for c in my_chunks}
        for p in payloads:
            for chunk, pos, raw, valid in _unpack(p):
                if chunk in raws:
                    raws[chunk][0][pos] = raw
                    raws[chunk][1][pos] = valid
        jobs = []                            # (chunk, code_type, pattern_type, score_btype, time, index, quality, soft)
        for chunk in my_chunks:
            idx, q, bt = self.final[chunk]
            for j in self.H.stage_jobs(self.key, idx, q, bt, raws[chunk][0], raws[chunk][1], self.rate):
                jobs.append((chunk,) + j)
        blobs = {c: b"" for c in my_chunks}
        if jobs:
            bits, err = self.ctx.viterbi([j[7] for j in jobs], [j[1] for j in jobs])
            parts = {c: [] for c in my_chunks}
            for i, (chunk, code_type, ptype, sbt, time, index, quality, soft) in enumerate(jobs):
                b = bits[i]
                parts[chunk].append(st.pack("<iddQfBBdH", 0, time, quality, index, float(err[i]), sbt, ptype, 1.0, len(b)) + b.tobytes())
            blobs = {c: b"".join(v) for c, v in parts.items()}
        return _pack(blobs)

    # -- stage 7: records -> the reference's merge (any rank can do it; run() leaves it to rank 0)
    def stage_merge(self, payloads) -> dict:
        blobs = [b"" for _ in self.plan]
        for p in payloads:
            for chunk, blob in _unpack(p).items():
                blobs[chunk] = blob
        return self.H.merge_chunks(blobs, [p[2] for p in self.plan], self.n_total / float(self.rate), [self.key], [""])

    def run(self, allgather) -> dict:
        if os.environ.get("AWM_TRACE"):
            return self._run_traced(allgather)
        pay = allgather(self.stage_peaks())
        retry = self.stage_select(pay)
        if retry:
            pay = allgather(self.stage_peaks(retry))
            self.stage_select(pay)
        self.stage_final(allgather(self.stage_refine()))
        pay = allgather(self.stage_decode())
        pay = allgather(self.stage_viterbi(pay))
        return self.stage_merge(pay) if self.rank == 0 else None       # the merged result lives on rank 0

    def _run_traced(self, allgather) -> dict:
        """run() with wall-clock stage times on stderr (development aid)"""
        import sys
        import time
        t = [time.perf_counter()]
        names = []

        def mark(name):
            t.append(time.perf_counter())
            names.append(name)
        p = self.stage_peaks(); mark("peaks")
        pay = allgather(p); mark("gather1")
        retry = self.stage_select(pay); mark("select")
        if retry:
            pay = allgather(self.stage_peaks(retry))
            self.stage_select(pay); mark("retry")
        r = self.stage_refine(); mark("refine")
        pay = allgather(r); mark("gather2")
        self.stage_final(pay); mark("final")
        d = self.stage_decode(); mark("decode")
        pay = allgather(d); mark("gather3")
        v = self.stage_viterbi(pay); mark("viterbi")
        pay = allgather(v); mark("gather4")
        doc = self.stage_merge(pay) if self.rank == 0 else None; mark("merge")
        if self.rank == 0:
            print("[trace] balanced get: " + " ".join("%s %.2f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t, t[1:])) + " total %.2f ms" % ((t[-1] - t[0]) * 1e3),
                  file=sys.stderr, flush=True)
        return doc
