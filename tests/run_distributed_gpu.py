"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29700 tests/run_distributed_gpu.py
Read-sharded LowHash0 (bucket all-to-all over NCCL) + replicated-marker alignment, compared on rank 0 with the CPU
oracle on the unsharded input. Prints 'DISTRIBUTED PARITY OK' on success."""
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
    p = synth.SynthParams(reads=1500, k=14, genome_markers=100000, n50_bases=15000, min_bases=8000, seed=71, palindromic_every=97)
    params = dict(m=4, hashFraction=0.01, minHashIterationCount=10, alignmentCandidatesPerRead=20.0, log2MinHashBucketCount=0,
                  minBucketSize=2, maxBucketSize=30, minFrequency=2)
    start, span, rev = synth.read_windows(p)
    bounds = D.balanced_read_ranges(span, world)
    rb, re = bounds[rank], bounds[rank + 1]
    ctx = capi.Context(local_rank)
    dm = capi.synth_generate_device(ctx, p, want_data7=False, read_begin=rb, read_end=re)
    total = torch.tensor([dm.marker_count], dtype=torch.int64, device="cuda")
    dist.all_reduce(total)
    ctx.set_markers_device(dm.toc, dm.kmer_ptr, dm.flags, keepalive=dm, read_begin=rb, read_end=re,
                           read_count_total=p.reads, total_marker_count=int(total.item()))
    stages = D.CudaStages(ctx, local_rank)
    cand, stats, info = D.lowhash0_sharded(stages, params, p.reads)
    allc = D.gather_candidates(cand)
    # alignment: replicate the markers, align the local slice, gather the record counts
    toc, gathered = D.all_gather_markers(ctx, local_rank, dm.toc)
    ctx2 = capi.Context(local_rank)
    ctx2.set_markers_device(toc, gathered.data_ptr(), dm.flags, keepalive=gathered)
    opts = dict(alignMethod=3, k=14, maxSkip=100, maxDrift=100, maxTrim=100, minAlignedMarkerCount=10, minAlignedFraction=0.1,
                downsamplingFactor=0.05, bandExtend=10, maxBand=1000)
    rec, ctoc, cdata, res = capi.compute_alignments(ctx2, cand, capi.make_align_options(**opts))
    recs = [None] * world if rank == 0 else None
    dist.gather_object((rec, ctoc, cdata), recs, dst=0)
    ok = True
    if rank == 0:
        d = synth.generate(p)
        oc, os_, _ = B.oracle_lowhash0(d["toc"], d["data"], d["flags"], B.LowHashParams(**params))
        ok = np.array_equal(allc, oc) and np.array_equal(stats.cpu().numpy().reshape(-1, 3).astype(np.uint64), os_)
        print(f"lowhash: {len(allc)} candidates, parity {'OK' if ok else 'MISMATCH'}; entries exchanged on rank 0: {info['entriesReceived']}")
        sample = oc[:: max(1, len(oc) // 1500)]
        orec, otoc, odata, _ = B.oracle_compute_alignments(d["toc"], d["kmer"], oc, B.make_align_options(**{k: v for k, v in opts.items() if k in B.ALIGN_DEFAULTS}), threads=32)
        grec = np.concatenate([r[0] for r in recs], axis=0)
        gdata = np.concatenate([r[2] for r in recs])
        ok2 = np.array_equal(grec, orec) and np.array_equal(gdata, odata)
        print(f"alignment: {len(grec)} stored alignments, parity {'OK' if ok2 else 'MISMATCH'}")
        ok = ok and ok2
        print("DISTRIBUTED PARITY OK" if ok else "DISTRIBUTED PARITY FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
