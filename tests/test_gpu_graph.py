"""Whole-step CUDA graph (efficientvlm_b200/graph.py) against the eager step it captured."""
import pytest
import torch

from tests.helpers import assert_close
from tests.test_gpu_models import _tiny_gd_models

pytestmark = pytest.mark.gpu


def _batch(B=6, L=9):
    g = torch.Generator().manual_seed(5)
    image = torch.randn(B, 3, 32, 32, generator=g)
    text_ids = torch.randint(1, 211, (B, L), generator=g)
    text_atts = torch.ones(B, L, dtype=torch.long)
    text_atts[1, 7:] = 0
    masked_pos = torch.stack([torch.randperm(L - 1, generator=g)[:3].sort().values + 1 for _ in range(B)])
    masked_ids = torch.gather(text_ids, 1, masked_pos)
    text_ids_masked = text_ids.clone().scatter_(1, masked_pos, 103)
    return [t.cuda() for t in (image, text_ids, text_atts, text_ids_masked, masked_pos, masked_ids)]


def _setup(train_mode, lr):
    from efficientvlm_b200.distill import gd_loss
    from efficientvlm_b200.optim import LinearWarmupDecay, create_optimizer
    student, teacher = _tiny_gd_models()
    student.cuda()
    teacher.cuda()
    student.train(train_mode)
    for p in teacher.parameters():
        p.requires_grad_(False)
    B = 6
    negs = (torch.roll(torch.arange(B), 1).cuda(), torch.roll(torch.arange(B), -2).cuda())
    sampler = lambda image_feat, text_feat, idx=None: negs  # noqa: E731
    student.sample_itm_negatives = sampler
    teacher.sample_itm_negatives = sampler
    opt = create_optimizer(dict(lr=lr, weight_decay=0.01, lr_mult=2), student, clip_grad_norm=1.0)
    sched = LinearWarmupDecay(opt, 10, 3)

    def device_step(*batch):
        so = student(*batch, output_attentions=True, output_hidden_states=True)
        with torch.no_grad():
            to = teacher(*batch, output_attentions=True, output_hidden_states=True)
        total, _ = gd_loss(so, to, 1.0)
        total.backward()
        opt.step(allreduce=False)
        opt.zero_grad()
        return total
    return student, opt, sched, device_step


def test_graphed_step_matches_eager_step():
    """6 optimizer steps with a moving learning rate (warm-up then decay): eager vs 2 eager warm-up + 4 graph replays."""
    from efficientvlm_b200.graph import GraphedTrainStep
    batch = _batch()
    s_e, opt_e, sched_e, step_e = _setup(False, 1e-3)
    losses_e = []
    for _ in range(6):
        losses_e.append(step_e(*batch).item())
        sched_e.step()
    s_g, opt_g, sched_g, step_g = _setup(False, 1e-3)
    graphed = GraphedTrainStep(step_g, batch, optimizers=[opt_g], warmup=2, host_fn=sched_g.step)
    assert graphed.captured_launches > 50
    losses_g = [graphed(*batch).item() for _ in range(4)]
    assert opt_g.state_step == 6 and sched_g.last_step == sched_e.last_step
    for a, b in zip(losses_e[2:], losses_g):
        assert abs(a - b) <= 2e-3 * abs(a), (losses_e, losses_g)
    assert losses_e[-1] < losses_e[0]          # it does train
    pe, pg = dict(s_e.named_parameters()), dict(s_g.named_parameters())
    for n in ("vision_encoder.encoder.layers.1.mlp.fc2.weight", "text_encoder.bert.encoder.layer.4.crossattention.self.key.weight",
              "itm_head.0.weight", "text_proj.bias"):
        assert_close(pg[n], pe[n], 2e-3, "param after 6 steps " + n)
    # the eager path still works on the same model after replays (bf16 shadows re-cast from the updated parameters)
    l_eager = step_g(*batch).item()
    sched_g.step()
    l_e7 = step_e(*batch).item()
    assert abs(l_eager - l_e7) <= 2e-3 * abs(l_e7)


def test_graph_replays_draw_fresh_dropout_masks():
    """lr = 0, dropout on: parameters never move, so the loss changes between replays only through the dropout masks, whose
    by-value seeds are baked into the graph — the device seed offset (evlm_rng_advance node) must move them."""
    from efficientvlm_b200.graph import GraphedTrainStep, reset_rng_offset
    batch = _batch()
    try:
        _, opt, sched, step = _setup(True, 0.0)
        graphed = GraphedTrainStep(step, batch, optimizers=[opt], warmup=1)
        losses = [graphed(*batch).item() for _ in range(4)]
        assert len({round(v, 6) for v in losses}) == 4, losses
        assert max(losses) - min(losses) < 0.2 * abs(losses[0])
    finally:
        reset_rng_offset()
        torch.cuda.synchronize()
