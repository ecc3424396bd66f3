"""Host-side control flow of the engine / training path with the kernel library STUBBED OUT (no GPU here): every C-ABI
call returns success without computing, so this exercises only what runs in Python — buffer shapes, dtypes, strides and the
argument checks of the ops wrappers, the autograd plumbing, tape staleness, the flat-parameter trainer — exactly the code
that otherwise meets a GPU for the first time on the box.  Numerics are covered by the `-m gpu` tests."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402


class _StubLib:
    """Returns MV_OK for every compute entry point; host-only size queries go to the real library."""

    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):
        if name in ("mv_lora_grads_workspace_bytes", "mv_loss_workspace_floats"):
            return getattr(self._real, name)

        def fn(*a):
            self.calls.append(name)
            return 0
        return fn


@pytest.fixture()
def stubbed(monkeypatch):
    from miphei_vit_b200 import lib, ops

    stub = _StubLib(lib.load())
    monkeypatch.setattr(ops, "_lib_for", lambda t: stub)
    monkeypatch.setattr(ops, "require_cuda", lambda device, what: None)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_sk_workspace", lambda device: None)
    return stub


def _model(out_chans=3, depth=2):
    from miphei_vit_b200.generators.mipheivit import get_vitmatte

    cfg = om.Config(img_size=128, embed_dim=128, depth=depth, num_heads=2, hidden=256, out_chans=out_chans)
    m = get_vitmatte("hoptimus0", cfg.img_size, cfg.out_chans, use_lora=True, embed_dim=cfg.embed_dim, depth=cfg.depth,
                     num_heads=cfg.num_heads, hidden=cfg.hidden)
    m.load_state_dict(om.init_state_dict(cfg, seed=3, perturb=True))
    return cfg, m


def test_eval_forward_control_flow(stubbed):
    cfg, m = _model()
    m.eval()
    m.engine.use_graphs = False
    x = torch.zeros(2, 3, 128, 128)
    with torch.no_grad():
        out = m(x)
        assert out.shape == (2, 3, 128, 128) and out.dtype == torch.float32
        assert m(x.half()).dtype == torch.float16
        assert m.engine.infer(x, out_dtype=torch.uint8).dtype == torch.uint8
        assert m.encoder(x).shape == (2, 128, 8, 8)
    assert "mv_gemm_bf16" in stubbed.calls and "mv_attn_fwd" in stubbed.calls
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 256, 256))


def test_plain_autograd_route_control_flow(stubbed):
    cfg, m = _model()
    m.train()
    x = torch.zeros(2, 3, 128, 128)
    pred = m(x)
    assert pred.requires_grad and pred.shape == (2, 3, 128, 128)
    stubbed.calls.clear()
    pred.sum().backward()
    named = dict(m.named_parameters())
    for n, p in named.items():
        if p.requires_grad:
            assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == torch.float32, n
        else:
            assert p.grad is None, n
    for k in ("mv_heads_bwd_algebra", "mv_gather_cast", "mv_attn_bwd", "mv_lora_grads", "mv_bn_relu_bwd"):
        assert k in stubbed.calls, k
    # stale tape: a second forward at the same batch size invalidates the first graph; a consumed graph cannot be re-used
    p1, p2 = m(x), m(x)
    with pytest.raises(RuntimeError, match="stale"):
        p1.sum().backward()
    p2.sum().backward()
    with pytest.raises(RuntimeError):
        p2.sum().backward()
    # under autocast the output follows the autocast dtype, gradients still arrive in fp32
    with torch.autocast("cpu", dtype=torch.bfloat16):
        pass  # (CUDA autocast state is what the engine reads; nothing to check on a CPU-only host)
    # eval-mode forward with grad enabled is not supported: loud, not silent
    m.eval()
    with pytest.raises(NotImplementedError):
        m(x)


@pytest.mark.parametrize("heads", [3, 16])
def test_trainer_step_control_flow(stubbed, heads):
    from miphei_vit_b200.trainer import Trainer

    cfg, m = _model(out_chans=heads, depth=3)
    tr = Trainer(m, marker_weights=torch.ones(heads), batch_size=2, total_steps=10, warmup_steps=2, use_graph=False)
    x, y = torch.zeros(2, 3, 128, 128), torch.zeros(2, heads, 128, 128)
    loss = tr.step(x, y)  # first step also packs the frozen weights
    assert loss.shape == (1,) and tr.step_count == 1
    stubbed.calls.clear()
    tr.step(x, y)
    first = list(stubbed.calls)
    stubbed.calls.clear()
    tr.step(x, y)
    assert stubbed.calls == first, "every step is the same kernel sequence (a CUDA graph can replay it)"
    for k in ("mv_lora_refresh", "mv_gather_cast", "mv_memset_async", "mv_add_i64", "mv_loss_fwd_bwd", "mv_grad_norm",
              "mv_adam_schedule", "mv_adam_clip_step_dev"):
        assert k in first, k
    assert first.count("mv_gather_cast") == 4      # three operand arenas + the gradient scatter
    assert first.count("mv_lora_refresh") == 1 and first.count("mv_add_i64") == 1
    stubbed.calls.clear()
    tr.step_autograd(x, y)
    assert tr.step_count == 4
    # parameters are views of the trainer's flat buffer; gradients of its flat gradient buffer
    for n, p in tr.order:
        assert p.grad is not None and p.grad.data_ptr() >= tr.gflat.data_ptr()
    # eval after training re-packs (weights generation moved)
    m.eval()
    m.engine.use_graphs = False
    with torch.no_grad():
        assert m(x).shape == (2, heads, 128, 128)
