import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import oracle_lib as O
from f2nerf_b200 import ops
T=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rng=np.random.default_rng(1)
for nh in (0,1):
    params=T((O.mlp_init(32,nh)*2).astype(np.float16))
    for n in (5000, 148*5*128+333, 148*5*128*2+333):
        x=T((rng.standard_normal((n,32))*0.5).astype(np.float16))
        out,hid=ops.mlp_fwd(x,params,nh,save_hidden=True,impl="tc")
        out0,hid0=ops.mlp_fwd(x,params,nh,save_hidden=True,impl="v0")
        for l in range(nh+1):
            a=hid[l].float().cpu().numpy(); b=hid0[l].float().cpu().numpy()
            bad=np.abs(a-b).max(1) > 0.02*np.abs(b).max()
            rows=np.nonzero(bad)[0]
            print(f"nh={nh} n={n} layer={l} bad rows={rows.size}", rows[:10], (rows//128)[:10] if rows.size else "", "tiles bad:", np.unique(rows//128).size)
        a=out.float().cpu().numpy(); b=out0.float().cpu().numpy()
        bad=np.abs(a-b).max(1) > 0.02*np.abs(b).max()
        print("   out bad rows", int(bad.sum()))
