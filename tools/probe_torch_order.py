#!/usr/bin/env python
"""Which FP32 summation order do the reference's torch-level projection ops use on THIS GPU?

The reference computes the projection with ATen / cuBLAS kernels (gs/renderer.py:381-419: one einsum for
project_pts, `rotmat @ rotmat^T`, an einsum for JW and two bmm for the covariance).  A fused kernel is bit-exact with
them only if it rounds in the same places.  This probe evaluates every stage with torch on the GPU and compares it,
bit for bit, with candidate orders emulated in float64 (a product of two floats is exact in double, so
float32(double(a)*double(b) + double(c)) is fma(a, b, c) up to a vanishing double-rounding rate).

  python tools/probe_torch_order.py [N]        -> JSON: match fraction of each candidate per stage
"""
import itertools
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gaussian_splatting_3d_b200 import synthetic as S  # noqa: E402
from oracle import ref_torch as R  # noqa: E402

dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
f32, f64 = torch.float32, torch.float64


def fma(a, b, c):
    return (a.to(f64) * b.to(f64) + c.to(f64)).to(f32)


def mul(a, b):
    return (a * b)


def dot3_candidates(a, b):
    """a, b: lists of three float32 tensors (terms a[k]*b[k]); -> {name: value}"""
    out = {}
    for perm in itertools.permutations(range(3)):
        i, j, k = perm
        tag = "".join(map(str, perm))
        out[f"fma_chain_{tag}"] = fma(a[k], b[k], fma(a[j], b[j], mul(a[i], b[i])))      # ((i) + j) + k, fused
        out[f"mul_add_{tag}"] = (mul(a[i], b[i]) + mul(a[j], b[j])) + mul(a[k], b[k])    # separate roundings
    return out


def report(name, ref, cands, res):
    r = {}
    refb = ref.contiguous().view(torch.int32)
    for k, v in cands.items():
        r[k] = float((v.contiguous().view(torch.int32) == refb).float().mean())
    best = max(r, key=r.get)
    exact = [k for k, v in r.items() if v == 1.0]
    res[name] = {"best": best, "match": round(r[best], 6), "n_exact": len(exact), "exact": exact[:4],
                 "top": {k: round(v, 5) for k, v in sorted(r.items(), key=lambda kv: -kv[1])[:3]}}


def main():
    res = {"N": N, "device": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "allow_tf32": torch.backends.cuda.matmul.allow_tf32}
    sc = S.make_scene("cfg2", seed=0, N=N)
    for pose_name, c2w in (("identity", sc["c2w"]), ("ring1", S.ring_cameras(8)[1])):
        c2w = c2w.to(dev)
        mean, qvec = sc["mean"].to(dev), sc["qvec"].to(dev)
        svec = torch.exp(sc["svec_before_activation"]).to(dev)
        # ---- project_pts: einsum("ij,bj->bi", W, p + d)
        d = -c2w[:3, 3]
        W = c2w[:3, :3].t()
        x = mean + d
        u = torch.einsum("ij,bj->bi", W, x)
        for i in range(3):
            a = [W[i, j].expand(N).contiguous() for j in range(3)]
            b = [x[:, j].contiguous() for j in range(3)]
            report(f"{pose_name}/project_pts[{i}]", u[:, i], dot3_candidates(a, b), res)
        # ---- sigma = A A^T
        A = R.qsvec2rotmat_batched(qvec, svec)
        sig = A @ A.transpose(-1, -2)
        for (i, j) in ((0, 0), (0, 1), (1, 2), (2, 2)):
            a = [A[:, i, k].contiguous() for k in range(3)]
            b = [A[:, j, k].contiguous() for k in range(3)]
            report(f"{pose_name}/sigma[{i}{j}]", sig[:, i, j], dot3_candidates(a, b), res)
        # ---- JW = einsum("bij,jk->bik", J, W)
        J = R.jacobian(u)
        JW = torch.einsum("bij,jk->bik", J, W)
        for (i, k) in ((0, 0), (0, 2), (1, 1), (2, 1)):
            a = [J[:, i, j].contiguous() for j in range(3)]
            b = [W[j, k].expand(N).contiguous() for j in range(3)]
            report(f"{pose_name}/JW[{i}{k}]", JW[:, i, k], dot3_candidates(a, b), res)
        # ---- X = bmm(JW, sigma); cov = bmm(X, JW^T)[:2,:2]
        X = torch.bmm(JW, sig)
        for (i, k) in ((0, 0), (0, 2), (1, 1)):
            a = [JW[:, i, j].contiguous() for j in range(3)]
            b = [sig[:, j, k].contiguous() for j in range(3)]
            report(f"{pose_name}/X[{i}{k}]", X[:, i, k], dot3_candidates(a, b), res)
        cov = torch.bmm(X, JW.transpose(-1, -2))
        for (i, k) in ((0, 0), (0, 1), (1, 0), (1, 1)):
            a = [X[:, i, j].contiguous() for j in range(3)]
            b = [JW[:, k, j].contiguous() for j in range(3)]
            report(f"{pose_name}/cov[{i}{k}]", cov[:, i, k], dot3_candidates(a, b), res)
        # ---- reductions inside elementwise-looking ops: torch.norm (jacobian) and F.normalize (quaternion)
        nrm = torch.norm(u, dim=-1)
        uu = [u[:, j].contiguous() for j in range(3)]
        cands = {f"sqrt({k})": torch.sqrt(v) for k, v in dot3_candidates(uu, uu).items()}
        cands["f64_then_round"] = torch.sqrt((u.to(f64) ** 2).sum(-1)).to(f32)
        cands["f64_sqrt32"] = torch.sqrt((u.to(f64) ** 2).sum(-1).to(f32))
        report(f"{pose_name}/norm(u)", nrm, cands, res)
        qn = torch.nn.functional.normalize(qvec, p=2.0, dim=-1, eps=1e-12)
        q = [qvec[:, j].contiguous() for j in range(4)]
        sq = [qq * qq for qq in q]
        n2 = {
            "mul_add_0123": ((sq[0] + sq[1]) + sq[2]) + sq[3],
            "mul_add_3210": ((sq[3] + sq[2]) + sq[1]) + sq[0],
            "fma_chain_0123": fma(q[3], q[3], fma(q[2], q[2], fma(q[1], q[1], sq[0]))),
            "fma_chain_3210": fma(q[0], q[0], fma(q[1], q[1], fma(q[2], q[2], sq[3]))),
            "pair_(01)(23)": (sq[0] + sq[1]) + (sq[2] + sq[3]),
            "pair_(02)(13)": (sq[0] + sq[2]) + (sq[1] + sq[3]),
            "pairfma_(01)(23)": fma(q[1], q[1], sq[0]) + fma(q[3], q[3], sq[2]),
            "pairfma_(02)(13)": fma(q[2], q[2], sq[0]) + fma(q[3], q[3], sq[1]),
            "f64": (qvec.to(f64) ** 2).sum(-1).to(f32),
        }
        qc = {k: q[0] / torch.clamp_min(torch.sqrt(v), 1e-12) for k, v in n2.items()}
        qc["f64_all"] = (qvec[:, 0].to(f64) / torch.sqrt((qvec.to(f64) ** 2).sum(-1))).to(f32)
        report(f"{pose_name}/normalize(q)[0]", qn[:, 0], qc, res)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
