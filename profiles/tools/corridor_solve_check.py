"""How accurate is ONE linear solve of the corridor graph (config 5 topology)?  Exports the assembled system
(H upper CSC, b) at the initial estimate, runs one GN iteration on the device, recovers dx from the estimates and
compares with scipy's sparse LU on the same system: relative residual and relative error in the 2-norm and the
H-norm."""
import sys, os, json
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
g = pkg.synth.c5_graph(n_poses=n_poses, n_pairs=n_poses // 10)
ctx = pkg.Context(0)
ctx.graph_load(g)
d = ctx.graph_export_system()
n = d["n"]
U = sp.csc_matrix((d["Ax"], d["Ai"], d["Ap"]), shape=(n, n))
H = U + sp.triu(U, 1).T
b = d["b"]
pe0, le0 = g.pose_est.copy(), g.lm_est.copy()
ctx.graph_prepare()
ctx.graph_iterate_async(1)
rc, chi2 = ctx.graph_finish()
pe1, le1 = ctx.graph_get_estimates()
# g2o Hessian order: free vertices by ascending id
fixed = set(int(v) for v in g.fixed_ids)
ids = [(int(v), 1, k) for k, v in enumerate(g.lm_ids)] + [(int(v), 0, k) for k, v in enumerate(g.pose_ids)]
ids.sort()
dx = []
for vid, is_lm, k in ids:
    if vid in fixed:
        continue
    if is_lm:
        dx += list(le1[k] - le0[k])
    else:
        dp = pe1[k] - pe0[k]
        dp[2] = (dp[2] + np.pi) % (2 * np.pi) - np.pi
        dx += list(dp)
dx = np.array(dx)
assert len(dx) == n, (len(dx), n)
out = {"n": int(n), "iterations": int(rc), "chi2": [float(v) for v in chi2]}
best = None
for sign in (1.0, -1.0):
    rhs = sign * b
    x_ref = spl.spsolve(H.tocsc(), rhs)
    r_gpu = H @ dx - rhs
    r_ref = H @ x_ref - rhs
    e = dx - x_ref
    res = {"rel_residual_gpu": float(np.linalg.norm(r_gpu) / np.linalg.norm(rhs)),
           "rel_residual_scipy": float(np.linalg.norm(r_ref) / np.linalg.norm(rhs)),
           "rel_err_2norm": float(np.linalg.norm(e) / np.linalg.norm(x_ref)),
           "rel_err_Hnorm": float(np.sqrt(abs(e @ (H @ e)) / abs(x_ref @ (H @ x_ref)))),
           "max_abs_err": float(np.max(np.abs(e))), "max_abs_dx": float(np.max(np.abs(x_ref)))}
    if best is None or res["rel_residual_gpu"] < best["rel_residual_gpu"]:
        best = res
out.update(best)
print(json.dumps(out))
