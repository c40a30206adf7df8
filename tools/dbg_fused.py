"""Debug helper: fused vs split on a small case, report where they differ (run on a GPU box)."""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beom_b200 import cases, model

def run(name, nsteps, **kw):
    c = cases.CASES[name](**kw)
    with tempfile.TemporaryDirectory() as d:
        blk = c.write(d)
        hm = model.HostModel.from_block(blk)
        res = {}
        for fused in (False, True):
            gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=fused))
            gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
            gm.advance(1, nsteps)
            res[fused] = gm.download_state() + gm.download_aux()
            print(name, "path", gm.path)
            gm.close()
        sub = hm.iarray("subc")
        for a, b, nm in zip(res[True], res[False], ("hlay", "u", "v", "h_u", "h_v", "rs_h", "dmdx", "dmdy")):
            a = np.asarray(a); b = np.asarray(b).reshape(a.shape)
            bad = np.argwhere(a != b)
            print("  %-5s mismatches %d of %d" % (nm, len(bad), a.size))
            if len(bad):
                pts = bad[:, -1] if nm in ("hlay", "u", "v", "h_u", "h_v") else bad[:, 1] if a.ndim == 3 else bad[:, -1]
                ii = sub[0][pts]; jj = sub[1][pts]
                print("     i range %d..%d  j range %d..%d ; first %s got %r want %r" % (ii.min(), ii.max(), jj.min(), jj.max(), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))
                print("     distinct i (first 20):", sorted(set(ii.tolist()))[:20], " distinct j (first 20):", sorted(set(jj.tolist()))[:20])
        hm.close()

if __name__ == "__main__":
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    run("synthetic_basin", ns, n=120, mm=60, nlay=2)
    run("stommel1948", ns, dl=250.0e3)
