"""Multi-GPU sharding of the ATRAC1 path: independent streams, then contiguous frame ranges.

Both directions are finite-window functions of their input: encoded
frame f of a channel depends on PCM from frame f-2 on, decoded frame f on sound units f-1 and
f.  A shard therefore needs a read-only halo and nothing else: ranks never exchange data and
write disjoint slices of the output (unit index = frame * n_ch + channel).  No collective.

plan() is pure host logic (every rank computes the same plan from the same arguments).
"""
from __future__ import annotations

from dataclasses import dataclass

ENC_HALO_FRAMES = 2  # carta1_encode_device: halo_frames is 0 (stream start) or >= 2
DEC_HALO_FRAMES = 1  # carta1_decode_device: halo_frames is 0 or >= 1
MIN_SHARD_FRAMES = 2


@dataclass(frozen=True)
class Shard:
    stream: int       # index into the caller's list of (multi-channel) streams
    begin: int        # first frame emitted
    end: int          # one past the last frame emitted
    enc_halo: int     # frames of PCM history to stage before `begin` for encode (0 or 2)
    dec_halo: int     # frames of sound units to stage before `begin` for decode (0 or 1)

    @property
    def frames(self) -> int:
        return self.end - self.begin

    def pcm_span(self):
        """[first, last) sample of every channel that the encode launch reads."""
        return (self.begin - self.enc_halo) * 512, self.end * 512

    def unit_span(self, n_ch: int):
        """[first, last) interleaved sound unit that the decode launch reads."""
        return (self.begin - self.dec_halo) * n_ch, self.end * n_ch


def plan(stream_frames, world: int):
    """Cut sum(stream_frames) frames into `world` contiguous, balanced spans of the stream-major
    frame order; a span that crosses a stream boundary becomes several shards.  Returns one list
    of Shard per rank.  Cuts never land on frame 1 of a stream (a 1-frame history would be a
    halo the kernels cannot tell from a stream start), and no shard is shorter than 2 frames
    unless its stream is."""
    if world < 1:
        raise ValueError("world must be >= 1")
    stream_frames = [int(n) for n in stream_frames]
    if any(n < 0 for n in stream_frames):
        raise ValueError("negative frame count")
    total = sum(stream_frames)
    starts, acc = [], 0
    for n in stream_frames:
        starts.append(acc)
        acc += n

    def legal(cut):  # snap a global frame index to a legal cut point
        for s, n in zip(starts, stream_frames):
            if s <= cut < s + n:
                local = cut - s
                if 0 < local < MIN_SHARD_FRAMES:
                    return s
                if n - local < MIN_SHARD_FRAMES:
                    return s + n
                return cut
        return total

    cuts = [0] + [legal((total * r) // world) for r in range(1, world)] + [total]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    out = []
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        shards = []
        for idx, (s, n) in enumerate(zip(starts, stream_frames)):
            a, b = max(lo, s), min(hi, s + n)
            if a < b:
                begin = a - s
                shards.append(Shard(idx, begin, b - s, ENC_HALO_FRAMES if begin else 0, DEC_HALO_FRAMES if begin else 0))
        out.append(shards)
    return out


def check_plan(shards_by_rank, stream_frames):
    """Raises unless the shards tile every stream exactly once."""
    seen = {i: [] for i in range(len(stream_frames))}
    for shards in shards_by_rank:
        for sh in shards:
            seen[sh.stream].append((sh.begin, sh.end))
    for i, n in enumerate(stream_frames):
        pos = 0
        for a, b in sorted(seen[i]):
            if a != pos or b <= a:
                raise AssertionError(f"stream {i}: gap or overlap at frame {pos} (next shard {a}..{b})")
            pos = b
        if pos != n:
            raise AssertionError(f"stream {i}: covered {pos} of {n} frames")
