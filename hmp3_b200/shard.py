"""Multi-GPU sharding of a batch of independent streams.

Streams never exchange data, so the only multi-GPU logic is WHO encodes WHAT: clips are sorted by duration
(longest first) and dealt round-robin so every rank gets about the same number of audio seconds; results are
gathered by stream index.  No collective touches the data path (torch.distributed is used only to gather the
small per-stream results, or not at all when every rank writes its own output files)."""
from typing import Callable, List, Sequence


def assign(durations: Sequence[float], world: int) -> List[List[int]]:
    """Stream indices per rank: longest-first, dealt round-robin in a serpentine order (0..w-1, w-1..0) so
    that the per-rank totals stay within one clip of each other."""
    order = sorted(range(len(durations)), key=lambda i: (-durations[i], i))
    shards = [[] for _ in range(world)]
    for k, i in enumerate(order):
        r = k % (2 * world)
        shards[r if r < world else 2 * world - 1 - r].append(i)
    return shards


def encode_sharded(controls, pcms, rank: int, world: int, encode: Callable, gather: Callable = None):
    """Encode this rank's shard with `encode(controls, pcms) -> list of byte arrays` (on the GPU:
    hmp3_b200.capi.encode_batch bound to the rank's device) and, if `gather` is given
    (e.g. torch.distributed.all_gather_object wrapped to return the list), return all streams' outputs in
    stream order on every rank; otherwise return {stream index: bytes} for the local shard."""
    mine = assign([p.shape[0] for p in pcms], world)[rank]
    outs = encode([controls[i] for i in mine], [pcms[i] for i in mine]) if mine else []
    local = dict(zip(mine, outs))
    if gather is None:
        return local
    merged = {}
    for part in gather(local):
        merged.update(part)
    return [merged[i] for i in range(len(pcms))]
