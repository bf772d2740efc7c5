"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29700 tests/run_distributed_gpu.py
Read-sharded LowHash0 + alignment through the library's own NCCL orchestration (csrc/dist.cu behind the C ABI:
shb_dist_init, shb_lowhash0_sharded, shb_compute_alignments_sharded), compared on rank 0 with the CPU oracle on the unsharded
input: candidates, ReadLowHashStatistics, AlignmentData, compressed bytes and the summed digests.
torch.distributed is only used to ship the NCCL id and to collect the ranks' outputs for the comparison.
Prints 'DISTRIBUTED PARITY OK' on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import bindings as B
    from shasta_b200 import capi, synth
    from shasta_b200 import distributed as D
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    reads = int(os.environ.get("SHB_DIST_TEST_READS", "1500"))
    p = synth.SynthParams(reads=reads, k=14, genome_markers=int(reads * 66), n50_bases=15000, min_bases=8000, seed=71, palindromic_every=97)
    params = dict(m=4, hashFraction=0.01, minHashIterationCount=10, alignmentCandidatesPerRead=20.0, log2MinHashBucketCount=0,
                  minBucketSize=2, maxBucketSize=30, minFrequency=2)
    start, span, rev = synth.read_windows(p)
    bounds = D.balanced_read_ranges(span, world)
    rb, re = bounds[rank], bounds[rank + 1]
    ctx = capi.Context(local_rank)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    ctx.dist_init(world, rank, uid.cpu().numpy().tobytes())
    dm = capi.synth_generate_device(ctx, p, want_data7=False, read_begin=rb, read_end=re)
    total = torch.tensor([dm.marker_count], dtype=torch.int64, device="cuda")
    dist.all_reduce(total)
    ctx.set_markers_device(dm.toc, dm.kmer_ptr, dm.flags, keepalive=dm, read_begin=rb, read_end=re,
                           read_count_total=p.reads, total_marker_count=int(total.item()))
    cand, stats, res = ctx.lowhash0_sharded(capi.make_lowhash_params(**params))
    opts = dict(alignMethod=3, k=14, maxSkip=100, maxDrift=100, maxTrim=100, minAlignedMarkerCount=10, minAlignedFraction=0.1,
                downsamplingFactor=0.05, bandExtend=10, maxBand=1000)
    rec, ctoc, cdata, ares = capi.compute_alignments_sharded(ctx, cand, capi.make_align_options(**opts))
    tm = ctx.dist_timing()
    outs = [None] * world if rank == 0 else None
    dist.gather_object((cand, rec, np.asarray(ctoc), np.asarray(cdata), res.candidateDigest, ares.alignmentDataDigest,
                        ares.compressedDigest, tm.asdict()), outs, dst=0)
    ok = True
    if rank == 0:
        M64 = (1 << 64) - 1
        d = synth.generate(p)
        oc, os_, _ = B.oracle_lowhash0(d["toc"], d["data"], d["flags"], B.LowHashParams(**params))
        allc = np.concatenate([o[0] for o in outs], axis=0)
        ok = np.array_equal(allc, oc) and np.array_equal(stats, os_)
        sizes = [len(o[0]) for o in outs]
        print(f"world {world}: lowhash {len(allc)} candidates (blocks {sizes}), parity {'OK' if ok else 'MISMATCH'}")
        orec, otoc, odata, _ = B.oracle_compute_alignments(d["toc"], d["kmer"], oc, B.make_align_options(**{k: v for k, v in opts.items() if k in B.ALIGN_DEFAULTS}), threads=32)
        grec = np.concatenate([o[1] for o in outs], axis=0)
        gdata = np.concatenate([o[3] for o in outs])
        ok2 = np.array_equal(grec, orec) and np.array_equal(gdata, odata)
        print(f"alignment: {len(grec)} stored alignments, parity {'OK' if ok2 else 'MISMATCH'}")
        dc = sum(o[4] for o in outs) & M64
        dr = sum(o[5] for o in outs) & M64
        db = sum(o[6] for o in outs) & M64
        ok3 = dc == capi.digest_candidates(oc) and dr == capi.digest_records(orec, 16) and db == capi.digest_compressed(orec, otoc, odata)
        print(f"digests summed over the ranks: candidates {dc:016x} alignment_data {dr:016x} compressed {db:016x} -> {'equal to the CPU path' if ok3 else 'MISMATCH'}")
        print("rank 0 timing (s):", {k: round(v, 4) if isinstance(v, float) else v for k, v in outs[0][7].items()})
        ok = ok and ok2 and ok3
        print("DISTRIBUTED PARITY OK" if ok else "DISTRIBUTED PARITY FAILED", flush=True)
    dist.barrier()
    ctx.dist_finalize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
