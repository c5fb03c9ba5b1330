"""One rank of tests/test_gpu_nccl.py (launched with torch.distributed.run, one process per GPU): the hash-sharded path over REAL NCCL
on fixtures of the reference; every rank writes its shard to <out>/<fixture>_rank<r>.npz for the test process to compare."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_dir, names = sys.argv[1], sys.argv[2].split(",")
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from conftest import load_golden
    from spades_for_blackbird_b200.host import binding as B
    from spades_for_blackbird_b200.host import distributed as D
    from spades_for_blackbird_b200.host import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")   # carries the NCCL id only; the data path is the library's own NCCL communicator
    ctx = B.Context(local)
    comm = D.nccl_comm(ctx, rank, world)
    for name in names:
        if name == "errorfree_k31":   # long chains: the whole-set extraction on rank 0
            reads, k, nb, tip = synth.codes_to_strings(synth.sample_pairs(synth.random_genome(6000, 77), 600, 120, 300, 0.0, 78)), 31, 20, None
        else:
            g = load_golden(name)
            reads, k, nb, tip = g["reads"], g["k"], g["buckets"], (int(g["tip_bound"]) if g["tip_bound"] >= 0 else None)
            if nb % world:
                nb *= world
        words, word_off, lens = O.pack_reads(reads)
        w, o, ln = D.slice_reads(words, word_off, lens, rank, world)
        streams = B.ReadStreams(ctx, w, o, ln)
        sh = B.construct_sharded(ctx, comm, streams, k, nb, tip_clip=tip is not None, tip_length_bound=tip or 0, gather_to=0)
        np.savez(os.path.join(out_dir, "%s_rank%d.npz" % (name, rank)), kpomers=sh.kpomers.final_kmers(), counts=sh.kpomers.counts(),
                 kmers=sh.kmers.final_kmers(), km_starts=sh.kmers.bucket_starts, kp_starts=sh.kpomers.bucket_starts, masks=sh.masks(), idx=sh.idx(),
                 index_bin=sh.index.serialize(), unitigs=np.array(sh.unitigs()), clipped=int(sh.info.clipped),
                 fallback=bool(sh.info.whole_set_fallback), nb=nb, bytes_sent=int(sh.info.bytes_sent))
        sh.free(); streams.free()
    comm.free()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
