"""GPU parity tests of the single kernels behind the C ABI (GEMM engines, attention step,
SCN cell step) against torch / the oracle."""
import pytest
import torch

from oracle import capdec_oracle as O
from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _ref_gemm(X, W, bias=None, addm=None):
    out = X.double().cpu() @ W.double().cpu().transpose(-1, -2)
    if bias is not None:
        out = out + bias.double().cpu()
    if addm is not None:
        out = out + addm.double().cpu()
    return out


SHAPES = [(32, 128, 64), (5, 37, 40), (32, 2048, 512), (70, 300, 1000), (200, 130, 72), (1, 8, 8),
          (1600, 1000, 512), (33, 4608, 512)]


@pytest.mark.parametrize("rows,N,K", SHAPES)
def test_gemm_simt_fp32(rows, N, K):
    from capdec import functional as CF
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + N)
    X = torch.randn(rows, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    addm = torch.randn(rows, N, device="cuda", generator=g)
    out = CF.gemm(X, W, bias=bias, addm=addm, precision="fp32")
    assert rel_err(out, _ref_gemm(X, W, bias, addm)) < 2e-6


def _pad_k(t, mult=8):
    """bf16 operand with a pitch that is a multiple of 8 elements (16 bytes) as TMA needs."""
    K = t.shape[-1]
    Kp = (K + mult - 1) // mult * mult
    buf = torch.zeros(*t.shape[:-1], Kp, dtype=torch.bfloat16, device=t.device)
    buf[..., :K] = t.to(torch.bfloat16)
    return buf[..., :K]


@pytest.mark.parametrize("rows,N,K", SHAPES)
def test_gemm_tc_bf16(rows, N, K):
    from capdec import functional as CF
    g = torch.Generator(device="cuda").manual_seed(rows * 11 + N)
    X = _pad_k(torch.randn(rows, K, device="cuda", generator=g))
    W = _pad_k(torch.randn(N, K, device="cuda", generator=g))
    bias = torch.randn(N, device="cuda", generator=g)
    addm = torch.randn(rows, N, device="cuda", generator=g)
    out = CF.gemm(X, W, bias=bias, addm=addm, precision="bf16")
    # inputs are exactly representable; only the fp32 accumulation order differs
    assert rel_err(out, _ref_gemm(X.float(), W.float(), bias, addm)) < 1e-5
    out_h = CF.gemm(X, W, bias=bias, precision="bf16", out_ft=True)
    assert out_h.dtype == torch.bfloat16
    assert rel_err(out_h.float(), _ref_gemm(X.float(), W.float(), bias)) < 1e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gemm_batched(precision):
    from capdec import functional as CF
    g = torch.Generator(device="cuda").manual_seed(5)
    ft = torch.float32 if precision == "fp32" else torch.bfloat16
    X = torch.randn(4, 24, 96, device="cuda", generator=g).to(ft)
    W = torch.randn(4, 64, 96, device="cuda", generator=g).to(ft)
    out = CF.gemm(X, W, precision=precision)
    assert out.shape == (4, 24, 64)
    assert rel_err(out, _ref_gemm(X.float(), W.float())) < 1e-5


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("rows,P,E,A", [(32, 196, 2048, 512), (3, 9, 40, 24), (7, 196, 128, 64),
                                        (130, 49, 256, 128)])
def test_attention_step_matches_oracle(precision, tol, rows, P, E, A):
    from capdec import functional as CF
    D = 32
    g = torch.Generator().manual_seed(rows + P)
    p = {"attention.encoder_att.weight": torch.randn(A, E, generator=g) / E ** 0.5,
         "attention.encoder_att.bias": torch.randn(A, generator=g) * 0.1,
         "attention.decoder_att.weight": torch.randn(A, D, generator=g) / D ** 0.5,
         "attention.decoder_att.bias": torch.randn(A, generator=g) * 0.1,
         "attention.full_att.weight": torch.randn(1, A, generator=g) / A ** 0.5 * 4,
         "attention.full_att.bias": torch.randn(1, generator=g)}
    enc = torch.randn(rows, P, E, generator=g).relu_()
    h = torch.randn(rows, D, generator=g)
    beta_pre = torch.randn(rows, E, generator=g)
    ft = torch.float32 if precision == "fp32" else torch.bfloat16
    if precision == "bf16":      # compare against the oracle on the rounded features
        enc = enc.to(ft).float()
    awe_ref, alpha_ref = O.soft_attention({k: v.double() for k, v in p.items()}, enc.double(), h.double())
    att1 = torch.nn.functional.linear(enc, p["attention.encoder_att.weight"],
                                      p["attention.encoder_att.bias"])
    att2 = torch.nn.functional.linear(h, p["attention.decoder_att.weight"],
                                      p["attention.decoder_att.bias"])
    g1 = torch.cat([att2, beta_pre], dim=1).cuda().contiguous()
    z, alpha, awe = CF.attention_step(att1.to(ft).cuda().contiguous(), enc.to(ft).cuda().contiguous(), g1, A,
                                      p["attention.full_att.weight"].reshape(-1).cuda().contiguous(),
                                      p["attention.full_att.bias"].cuda(), precision=precision)
    assert rel_err(alpha, alpha_ref) < tol
    assert rel_err(awe, awe_ref) < tol
    z_ref = torch.sigmoid(beta_pre.double()) * awe_ref
    assert rel_err(z.float(), z_ref) < max(tol, 1e-5 if precision == "fp32" else 1e-2)
    assert (alpha.sum(dim=1) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("rows,X,D,F,S", [(32, 2560, 512, 512, 1000), (5, 56, 32, 24, 12)])
def test_scn_cell_step_matches_oracle(precision, tol, rows, X, D, F, S):
    from capdec import functional as CF
    g = torch.Generator().manual_seed(rows)
    bound = 1.0 / D ** 0.5
    names = ["weight_ia", "weight_ib", "weight_ic", "weight_ha", "weight_hb", "weight_hc", "bias_ih",
             "bias_hh"]
    shapes = [(X, 4 * F), (S, 4 * F), (D, 4 * F), (D, 4 * F), (S, 4 * F), (D, 4 * F), (4 * D,), (4 * D,)]
    p = {"c." + n: (torch.rand(s, generator=g) * 2 - 1) * bound for n, s in zip(names, shapes)}
    x = torch.randn(rows, X, generator=g) * 0.5
    s = torch.rand(rows, S, generator=g)
    h = torch.randn(rows, D, generator=g) * 0.5
    c = torch.randn(rows, D, generator=g)
    pd = {k: v.double() for k, v in p.items()}
    h_ref, c_ref = O.scn_cell(pd, "c.", x.double(), s.double(), h.double(), c.double())
    weights = [p["c." + n].cuda() for n in names]
    h_out, c_out = CF.scn_cell_step(weights, x.cuda(), s.cuda(), h.cuda(), c.cuda(), precision=precision)
    assert rel_err(h_out, h_ref) < tol
    assert rel_err(c_out, c_ref) < tol


def test_ops_refuse_cpu_tensors():
    from capdec import functional as CF
    from capdec._lib import CapdecError
    with pytest.raises(CapdecError):
        CF.gemm(torch.zeros(4, 8), torch.zeros(4, 8), precision="fp32")


@pytest.mark.parametrize("rows,N,K,splitk", [(32, 2048, 2048, -1), (32, 512, 2560, 16), (32, 4608, 512, -1),
                                             (20, 300, 1000, 3), (64, 512, 2048, -1)])
def test_gemm_tc_splitk_atomics(rows, N, K, splitk):
    """split-K slices reduced with fp32 atomics into a pre-initialised output (in-place addend)."""
    from capdec import functional as CF
    g = torch.Generator(device="cuda").manual_seed(rows + N + K)
    X = _pad_k(torch.randn(rows, K, device="cuda", generator=g))
    W = _pad_k(torch.randn(N, K, device="cuda", generator=g))
    bias = torch.randn(N, device="cuda", generator=g)
    out = CF.gemm(X, W, bias=bias, precision="bf16", splitk=splitk)
    assert rel_err(out, _ref_gemm(X.float(), W.float(), bias)) < 1e-5
    acc = torch.randn(rows, N, device="cuda", generator=g)
    ref = _ref_gemm(X.float(), W.float(), None, acc)
    CF.gemm(X, W, addm=acc, out=acc, precision="bf16", splitk=splitk)      # acc += X W^T
    assert rel_err(acc, ref) < 1e-5


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("rows,P,E,A", [(32, 196, 2048, 512), (4, 196, 2048, 512), (3, 9, 40, 24),
                                        (130, 49, 256, 128)])
def test_attention_bwd_step_matches_autograd(precision, tol, rows, P, E, A):
    """Backward kernel vs torch autograd (fp64) of the same math.  The pre-activations are kept
    away from the relu kink (|att1+att2| > margin) so that no mask can flip between precisions."""
    from capdec import functional as CF
    g = torch.Generator().manual_seed(rows * 3 + P)
    ft = torch.float32 if precision == "fp32" else torch.bfloat16
    att2 = torch.randn(rows, A, generator=g) * 0.3
    att1 = torch.randn(rows, P, A, generator=g)
    att1 = (att1.sign() * (att1.abs() + 0.05)).to(ft).float()          # representable in ft
    pre = att1 + att2.unsqueeze(1)
    att1 = torch.where(pre.abs() < 0.02, att1 + 0.1 * pre.sign() + 0.1 * (pre == 0), att1).to(ft).float()
    enc = torch.randn(rows, P, E, generator=g).relu_().to(ft).float()
    beta_pre = torch.randn(rows, E, generator=g)
    w_f = torch.randn(A, generator=g) / A ** 0.5 * 3
    b_f = torch.randn(1, generator=g)
    dz = torch.randn(rows, E, generator=g)
    dal = torch.randn(rows, P, generator=g) * 0.1
    # fp64 autograd reference
    a1 = att1.double().requires_grad_(True)
    a2 = att2.double().requires_grad_(True)
    bp = beta_pre.double().requires_grad_(True)
    wf = w_f.double().requires_grad_(True)
    bf = b_f.double().requires_grad_(True)
    e = torch.relu(a1 + a2.unsqueeze(1)) @ wf + bf
    alpha = torch.softmax(e, dim=1)
    awe = (enc.double() * alpha.unsqueeze(2)).sum(1)
    z = torch.sigmoid(bp) * awe
    ((z * dz.double()).sum() + (alpha * dal.double()).sum()).backward()
    g1 = torch.cat([att2, beta_pre], dim=1).cuda().contiguous()
    dbeta, datt2, dAtt1, dwf, dbf = CF.attention_bwd_step(
        att1.to(ft).cuda().contiguous(), enc.to(ft).cuda().contiguous(), g1, A, w_f.cuda(),
        alpha.float().cuda().contiguous(), dz.cuda(), awe.float().cuda().contiguous(),
        dalpha_ext=dal.cuda(), precision=precision)
    assert rel_err(dbeta.float(), bp.grad) < tol
    assert rel_err(datt2.float(), a2.grad) < tol
    assert rel_err(dAtt1, a1.grad) < tol
    assert rel_err(dwf.sum(0), wf.grad) < tol
    assert abs(dbf.sum().item() - bf.grad.item()) < 1e-4


@pytest.mark.parametrize("rows,N,K", [(512, 512, 1600), (10000, 512, 1600), (64, 200, 96), (2560, 2048, 1600),
                                      (100, 56, 37 * 8)])
def test_gemm_tn_matches_torch(rows, N, K):
    """Transposed-operand (MN-major tcgen05) weight-gradient product against fp32 torch on the same bf16 values."""
    from capdec import functional as CF
    g = torch.Generator(device="cuda").manual_seed(rows + N + K)
    ldx, ldw = (rows + 7) // 8 * 8, (N + 7) // 8 * 8
    XT = torch.randn(K, ldx, device="cuda", generator=g).to(torch.bfloat16)[:, :rows]
    WT = torch.randn(K, ldw, device="cuda", generator=g).to(torch.bfloat16)[:, :N]
    out = CF.gemm_tn(XT, WT)
    ref = XT.float().t() @ WT.float()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err
