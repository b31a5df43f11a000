"""Generates tests/golden/dualgnn_ico3.npz from the CPU oracle (the reference cannot be imported
here, SURVEY.md 8c, so these vectors pin the oracle against drift, not against upstream).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import util  # noqa: E402


def main():
    torch.set_num_threads(1)
    (dv, df), mesh_n, mesh_o = util.oracle_inputs(3, seed=0)
    net = util.oracle_net(0)
    util.set_perm_fn(net, 1234)
    net.record = True
    dv_in, df_in = dv.clone(), df.clone()
    with torch.no_grad():
        vp, nrm, _ = net([dv, df])
    out = dict(points_noisy=mesh_n.points, faces=mesh_n.fv,
               x_v=dv_in.x.numpy(), ei_v=dv_in.edge_index.numpy(), w_v=dv_in.edge_weight.numpy(), y_v=dv_in.y.numpy(),
               x_f=df_in.x.numpy(), ei_f=df_in.edge_index.numpy(), w_f=df_in.edge_weight.numpy(), y_f=df_in.y.numpy(),
               vert_p=vp.numpy(), norm_p=nrm.numpy(), xf12=net.taps["xf12"].numpy(),
               g_v=net.taps["g_v"].numpy(), g_f=net.taps["g_f"].numpy())
    for name, pl in zip(("v1", "v2", "f1", "f2"), util.poolings(net)):
        for s, (ei, w, perm, raw) in enumerate(pl.trace):
            out[f"label_{name}_{s}"] = raw.numpy()
            out[f"perm_{name}_{s}"] = perm.numpy()
        out[f"unpool_{name}"] = pl.unpooling_indices.numpy()
    for g in ("v", "f"):
        for k, v in net.taps[g].items():
            out[f"tap_{g}_{k}"] = v.numpy()
    path = os.path.join(util.GOLDEN, "dualgnn_ico3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
