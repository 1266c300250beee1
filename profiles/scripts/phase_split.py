"""L2 behaviour of the upper Occ levels: the same kernels on reads of 8/16/24 bp (the first LF steps only) and 100 bp,
2 Gbp index.  Run under ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum,
lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum -k regex:fm_search ."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200"); L = pkg.lib()
n, nq = 2_000_000_000, 4_000_000
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free(); idx.fuse()
stream = torch.cuda.current_stream().cuda_stream
for length in (8, 16, 24, 32, 100):
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    wpq = L.fmgpu_words_per_query(length)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
    for name, v in (("coop", pkg.variant(pkg.MODE_COOP, 1, 256)), ("fused", pkg.variant(pkg.MODE_FUSED, 2))):
        for _ in range(2):      # second launch of each pair is the warm one
            pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search")
        torch.cuda.synchronize()
        print(json.dumps({"len": length, "kernel": name}), flush=True)
