"""CPU, world_size = 2 (gloo): the data-parallel plumbing of simseg_b200.dist — row gather, gradient reduce-scatter
(GatherLayer.backward semantics, utils/dist.py:348-354 of the reference) and the flat-buffer mean all-reduce — checked
against the fixture the REFERENCE produced on 2 gloo ranks (tests/golden/global_reduce.npz)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "global_reduce.npz")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import simseg_oracle as O            # checker for the loss math (the CUDA loss kernels need a GPU)
        from simseg_b200 import dist as sdist
        z = np.load(GOLD)
        img, txt = torch.tensor(z["gr_img"]), torch.tensor(z["gr_txt"])
        b = img.shape[0] // world
        li, lt = img[rank * b:(rank + 1) * b].clone(), txt[rank * b:(rank + 1) * b].clone()
        out = {}
        # ---- forward gather: rank-major row order, identical on every rank
        tg = sdist.all_gather_rows(lt, sdist.WORLD)
        ig = sdist.all_gather_rows(li, sdist.WORLD)
        out["gather_ok"] = bool(torch.equal(tg, txt) and torch.equal(ig, img))
        # ---- loss of this rank and gradients: local-row part + reduce-scatter of the gathered-operand part
        li_, lt_ = li.clone().requires_grad_(True), lt.clone().requires_grad_(True)
        ig_, tg_ = ig.clone().requires_grad_(True), tg.clone().requires_grad_(True)
        loss, _, _ = O.clip_loss(li_, lt_, ig_, tg_, torch.tensor(0.02), rank)
        loss.backward()
        d_img = li_.grad + sdist.reduce_scatter_rows(ig_.grad, rank, b, sdist.WORLD)
        d_txt = lt_.grad + sdist.reduce_scatter_rows(tg_.grad, rank, b, sdist.WORLD)
        out["loss_err"] = abs(loss.item() - float(z[f"gr_loss_{rank}"]))
        out["dimg_err"] = (d_img - torch.tensor(z[f"gr_dimg_{rank}"])).abs().max().item()
        out["dtxt_err"] = (d_txt - torch.tensor(z[f"gr_dtxt_{rank}"])).abs().max().item()
        # ---- flat gradient buffer: mean all-reduce, .grad views stay attached
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))]
        fg = sdist.FlatGrads(ps)
        fg.zero()
        ps[0].grad.add_(float(rank + 1))
        ps[1].grad.add_(10.0 * (rank + 1))
        fg.all_reduce_async()
        fg.wait()
        out["flat_ok"] = bool(torch.allclose(ps[0].grad, torch.full((3, 5), 1.5)) and
                              torch.allclose(ps[1].grad, torch.full((7,), 15.0)) and
                              ps[0].grad.data_ptr() == fg.flat.data_ptr())
        ps[0].grad = None                                  # an optimizer with set_to_none=True
        fg.zero()
        out["reattach_ok"] = ps[0].grad is not None and ps[0].grad.data_ptr() == fg.flat.data_ptr()
        out["rank_world"] = (sdist.rank(), sdist.world_size())
        # ---- BERT dropout stream: same {seed, step} buffer on every rank (a DDP wrap broadcasts buffers), rank-specific key
        from simseg_b200.config import load_cfg
        from simseg_b200.pipeline import HuggingFaceModel, _Shared
        cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False"])
        real = HuggingFaceModel.__init__

        def light(self, cfg_, shared, **kw):               # the stream logic without allocating BERT-base on two CPU ranks
            torch.nn.Module.__init__(self)
            self.hidden_dropout_prob = self.attention_probs_dropout_prob = 0.1
            self.register_buffer("drop_rng", torch.zeros(2, dtype=torch.int64), persistent=False)
            self._drop_seeded = False
        HuggingFaceModel.__init__ = light
        try:
            hf = HuggingFaceModel(cfg, _Shared())
        finally:
            HuggingFaceModel.__init__ = real
        hf.seed_dropout(1234, 7)
        d = hf._dropout()
        out["drop_buf"] = hf.drop_rng.tolist()
        out["drop_key"] = d.rng.tolist()
        hf.eval()
        out["drop_eval_none"] = hf._dropout() is None
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_reduce_scatter_and_flat_grads():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=240) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for r in range(2):
        o = res[r]
        assert o["gather_ok"] and o["flat_ok"] and o["reattach_ok"]
        assert o["rank_world"] == (r, 2)
        assert o["loss_err"] < 1e-5 and o["dimg_err"] < 1e-5 and o["dtxt_err"] < 1e-5, o
        assert o["drop_buf"] == [1234, 8] and o["drop_eval_none"]           # the forward bumped step; eval draws nothing
        assert o["drop_key"] == [1234 + ((r * 0x9E3779B97F4A7C15) & 0x3FFFFFFFFFFFFFFF), 8]
    assert res[0]["drop_key"] != res[1]["drop_key"]


def test_single_process_paths_are_identity():
    from simseg_b200 import dist as sdist
    x = torch.randn(4, 8)
    assert sdist.all_gather_rows(x, None) is x
    assert sdist.reduce_scatter_rows(x, 0, 4, None) is x
    assert sdist.rank() == 0 and sdist.world_size() == 1


def test_qkv_gradients_are_one_matrix_under_flatgrads():
    """``towers.qkv_adjacent_order`` + ``dist.FlatGrads``: a BERT layer's query / key / value weight gradients form ONE
    contiguous [3D, D] matrix (one wgrad GEMM in ``bert_backward``), the biases one [3D] vector; every parameter appears
    exactly once and the packed view aliases the per-parameter ``.grad`` tensors."""
    import torch
    from simseg_b200 import towers
    from simseg_b200.dist import FlatGrads
    from simseg_b200.pipeline import BertModel
    bert = BertModel(vocab=100, dim=64, heads=1, ffn=128, depth=2, max_pos=16)
    params = list(bert.parameters())
    order = towers.qkv_adjacent_order(bert.named_parameters())
    assert len(order) == len(params) and {id(p) for p in order} == {id(p) for p in params}
    FlatGrads(order)
    for layer in bert.encoder.layer:
        qp = towers._qkv_params(layer)
        w = towers._adjacent([p.weight.grad for p in qp])
        b = towers._adjacent([p.bias.grad for p in qp])
        assert w is not None and w.shape == (192, 64) and b is not None and b.shape == (192,)
        w[64:128].fill_(2.0)
        b[128:].fill_(3.0)
        assert qp[1].weight.grad.eq(2.0).all() and qp[0].weight.grad.eq(0.0).all() and qp[2].weight.grad.eq(0.0).all()
        assert qp[2].bias.grad.eq(3.0).all() and qp[1].bias.grad.eq(0.0).all()
    # module order (weight, bias interleaved) is not adjacent: bert_backward then keeps three GEMMs
    FlatGrads(params)
    assert towers._adjacent([p.weight.grad for p in towers._qkv_params(bert.encoder.layer[0])]) is None
