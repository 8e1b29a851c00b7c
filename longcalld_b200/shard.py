"""Region sharding across the GPUs of one box and the final result gather.

The reference processes a contig as independent 500 kb region chunks (LONGCALLD_BAM_CHUNK_REG_SIZE, reference
src/bam_utils.h:10) that only meet again in stitch_var_main / make_var_main / write_var_to_vcf
(src/call_var_main.c:776-803).  One process drives one GPU; chunks are dealt in contiguous blocks round-robin
(neighbouring chunks share overlap reads, so a block keeps them on one device), every rank runs the hot path on its
own chunks with no data-path collective, and the per-chunk result blobs travel once, at the end of a contig, to the
rank that stitches and writes the VCF.  The gather is a torch.distributed collective: NCCL over NVLink/NVSwitch when the
payload lives on the GPUs, gloo in the CPU tests.
"""
import numpy as np

CHUNK_BLOCK = 8          # consecutive region chunks that stay on one rank (4 Mb of reference)


def deal_chunks(n_chunks, world, block=CHUNK_BLOCK):
    """rank -> sorted chunk indices: contiguous blocks of `block` chunks dealt round-robin."""
    owner = (np.arange(n_chunks) // block) % max(world, 1)
    return [np.nonzero(owner == r)[0] for r in range(world)]


def owner_of(n_chunks, world, block=CHUNK_BLOCK):
    return (np.arange(n_chunks) // block) % max(world, 1)


def pack_blobs(blobs):
    """[bytes-like per chunk] -> (uint8 payload, int64 lengths)"""
    lens = np.fromiter((len(b) for b in blobs), dtype=np.int64, count=len(blobs))
    payload = np.frombuffer(b"".join(bytes(b) for b in blobs), dtype=np.uint8) if len(blobs) else np.zeros(0, np.uint8)
    return payload, lens


def gather_chunk_results(my_chunks, my_blobs, n_chunks, dst=0, device=None):
    """Gather the variable-length result blob of every region chunk to rank `dst`, returned there in chunk order
    (what stitch_var_main walks); other ranks get None.  Two collectives: the lengths (all_gather, fixed size) and
    the concatenated payloads padded to the longest rank (gather)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    owner = owner_of(n_chunks, world)
    assert len(my_chunks) == len(my_blobs) and all(owner[c] == rank for c in my_chunks)
    payload, lens = pack_blobs(my_blobs)
    lens_all = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    if len(my_chunks):
        lens_all[torch.as_tensor(np.asarray(my_chunks), device=dev)] = torch.as_tensor(lens, device=dev)
    dist.all_reduce(lens_all, op=dist.ReduceOp.SUM)                       # every chunk has exactly one owner
    lens_np = lens_all.cpu().numpy()
    per_rank = np.array([int(lens_np[owner == r].sum()) for r in range(world)], dtype=np.int64)
    width = int(per_rank.max()) if world else 0
    send = torch.zeros(max(width, 1), dtype=torch.uint8, device=dev)
    if payload.size:
        send[:payload.size] = torch.as_tensor(payload.copy(), device=dev)
    recv = [torch.zeros(max(width, 1), dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    out = [None] * n_chunks
    for r in range(world):
        buf = recv[r].cpu().numpy()
        off = 0
        for c in np.nonzero(owner == r)[0]:
            out[c] = buf[off:off + lens_np[c]].tobytes()
            off += int(lens_np[c])
    return out
