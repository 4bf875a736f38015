"""Debug aid: runs the error paths of the compiled extension one per subprocess, with a native backtrace on SIGSEGV."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    "knn_ok": "d,i=ext.knn(r,q,5); torch.cuda.synchronize(); print('ok', d.shape)",
    "knn_err": "\ntry:\n    ext.knn(r,q,101)\nexcept RuntimeError as e:\n    print('raised', str(e)[:120])",
    "graph_err": "\ntry:\n    ext.knn_graph(x,20,4000)\nexcept RuntimeError as e:\n    print('raised', str(e)[:120])",
    "op_err": "\ntry:\n    torch.ops.gcanet_b200_native.knn(r,q,101,1)\nexcept RuntimeError as e:\n    print('raised', str(e)[:120])",
}
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import torch
    ctypes.CDLL(os.path.join(ROOT, "tools", "bin", "libsegv_trace.so"))
    from gcanet_b200 import native_ext
    ext = native_ext.load()
    r, q, x = torch.randn(3, 100, device="cuda"), torch.randn(3, 50, device="cuda"), torch.randn(1, 3, 2000, device="cuda")
    exec(CASES[sys.argv[1]])
    sys.exit(0)
for name in CASES:
    p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True)
    print(f"== {name}: rc={p.returncode}\n{p.stdout[-400:]}\n{p.stderr[-3000:]}", flush=True)
