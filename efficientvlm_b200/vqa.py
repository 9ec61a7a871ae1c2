"""VQA task models on the B200 hot path — drop-in for `efficient_models/model_generation.py::EffXVLMForVQA` (L0-gated student,
BASELINE configs 3 and 5) and `models/model_generation.py::XVLMForVQA` (un-gated teacher), plus the loss assembly of the
`Eff_VQA.py` training step.

Same constructor config keys, `forward(image, quesiton, answer, k, weights, train, output_attentions, output_hidden_states,
stop_prune)` signature [sic: the reference spells the argument `quesiton`], result dict layout, `rank_answer`, `tile`.

What changes underneath (results identical, tests/test_gpu_models.py::test_vqa_*):
* the per-question Python loops that replicate question states for their answers (model_generation.py:134-139, 262-281) become
  a row -> question index handed to the decoder's cross-attention (`encoder_batch_index`): the K/V projections of the question
  states run once per question instead of once per answer candidate (x128 at k_test = 128), and nothing is replicated in HBM;
* `rank_answer` is free of host synchronisation: top-k, candidate gather and re-ranking stay on the device;
* all KD terms of a step are ONE multi-pair MSE launch (`vqa_kd_losses`).
"""
import copy

import torch

from . import checkpoint, ops
from .distill import _kd_pairs, get_cor_teacher
from .eff_bert import BertLMHeadModel
from .l0_module import VQAL0Module
from .xvlm import XVLMBase, load_pretrained


def tile(x, dim, n_tile):
    """model_generation.py:15-21: every slice along `dim` repeated n_tile times, consecutively."""
    return torch.repeat_interleave(x, n_tile, dim=dim)


def _rows_of(k, device):
    """answer row -> question index for `k` answers per question (model_generation.py:134-139); k: list / tensor of ints or int."""
    if torch.is_tensor(k):
        k = k.tolist()
    idx = [b for b, n in enumerate(k) for _ in range(int(n))]
    return torch.tensor(idx, dtype=torch.int32).to(device, non_blocking=True)


class _Ungated:
    gated = False


class XVLMForVQA(XVLMBase):
    """models/model_generation.py:228-443 — un-gated VQA model (the distillation teacher of Eff_VQA.py)."""
    gated = False

    def __init__(self, config):
        super().__init__(config, load_vision_params=False, load_text_params=False, use_contrastive_loss=False, use_matching_loss=False,
                         use_mlm_loss=False, use_bbox_loss=False, config_text=None)
        assert isinstance(config["pad_token_id"], int)
        self.pad_token_id = config["pad_token_id"]
        config_enc = self.text_encoder.config
        self.num_text_layers = config_enc.fusion_layer
        self.num_cross_layers = config_enc.num_hidden_layers - config_enc.fusion_layer
        assert config["num_dec_layers"] == self.num_cross_layers, "initialization not implemented"
        config_dec = copy.deepcopy(config_enc)
        config_dec.encoder_width = config_enc.hidden_size
        config_dec.fusion_layer = 0  # start index
        config_dec.num_hidden_layers = config["num_dec_layers"]
        self.cross_encoder_width = config_enc.encoder_width  # i.e. vision_width
        self.dec_encoder_width = config_enc.hidden_size
        self.text_decoder = BertLMHeadModel(config=config_dec)
        if self.gated:
            self.l0_module = VQAL0Module(config, target_sparsity=config["sparsity"])
        if self.dec_encoder_width != self.cross_encoder_width:
            self.init_params = ["text_decoder." + n for n, _ in self.text_decoder.named_parameters()
                                if ("crossattention.self.key" in n) or ("crossattention.self.value" in n)]
        else:
            self.init_params = []

    def load_pretrained(self, ckpt_rpath, config, is_eval=False):
        """model_generation.py:52-96: the decoder is initialised from the fusion layers of the pre-trained text encoder."""
        state_dict = load_pretrained(ckpt_rpath, config, is_eval=True) if is_eval else load_pretrained(ckpt_rpath, config, load_text=False)
        if not is_eval:
            checkpoint.remap_keys(state_dict, checkpoint.vqa_decoder_rule(self.num_text_layers, self.dec_encoder_width != self.cross_encoder_width))
        checkpoint.load_into(self, state_dict, ckpt_rpath)

    # ------------------------------------------------------------------------------------------------------------------
    def _gates(self, train, stop_prune):
        """Eight gate tensors in the reference's routing (model_generation.py:98-113, 190-201) or Nones for the teacher."""
        if not self.gated:
            return dict(vision_head=None, vision_mlp=None, enc_head=None, enc_mlp=None, dec_head=None, dec_mlp=None)
        if train and not stop_prune:
            zs = self.l0_module.forward(training=True)
        else:
            with torch.no_grad():
                zs = self.l0_module.forward(training=False)
        return dict(vision_head=zs["vision_head_z"], vision_mlp=zs["vision_intermediate_z"],
                    enc_head=torch.cat((zs["text_head_z"], zs["cross_head_z"]), dim=0),
                    enc_mlp=torch.cat((zs["text_intermediate_z"], zs["cross_intermediate_z"]), dim=0),
                    dec_head=zs["decoder_head_z"], dec_mlp=zs["decoder_intermediate_z"])

    def forward(self, image, quesiton, answer=None, k=None, weights=None, train=True, output_attentions=None, output_hidden_states=None,
                stop_prune=False, answer_rows=None, use_gates=True):
        """answer_rows (extension, int32 device tensor [n_answers]): the answer row -> question index that `k` (answers per
        question) expands to; a caller that keeps it on the device avoids the per-step host list and copy.
        use_gates=False (extension): run without L0 gates — for a model whose masks were materialised (prune.materialize)."""
        z = self._gates(train, stop_prune) if use_gates else XVLMForVQA._gates(_Ungated, train, stop_prune)
        kd = bool(output_attentions) and train
        if kd:
            image_embeds, image_hidden_states, image_attentions = self.vision_encoder(
                image, output_attentions=output_attentions, output_hidden_states=output_hidden_states, head_z=z["vision_head"],
                mlp_z=z["vision_mlp"])
        else:
            image_embeds = self.vision_encoder(image, head_z=z["vision_head"], mlp_z=z["vision_mlp"])[0]
        # image_atts is all ones (model_generation.py:117): "no mask" on this path
        oa, oh = (output_attentions, output_hidden_states) if kd else (None, None)
        question_output = self.text_encoder(quesiton.input_ids, attention_mask=quesiton.attention_mask, encoder_hidden_states=image_embeds,
                                            encoder_attention_mask=None, return_dict=True, output_attentions=oa, output_hidden_states=oh,
                                            head_z=z["enc_head"], mlp_z=z["enc_mlp"])
        if not train:
            topk_ids, topk_probs, _ = self.rank_answer(question_output.last_hidden_state, quesiton.attention_mask, answer.input_ids,
                                                       answer.attention_mask, k, [z["dec_head"], z["dec_mlp"]])
            return topk_ids, topk_probs
        # k answers per question: answer row r reads question rows_of[r] through the cross-attention index
        rows_of = answer_rows if answer_rows is not None else _rows_of(k, image.device)
        answer_targets = answer.input_ids.masked_fill(answer.input_ids == self.pad_token_id, -100)
        answer_output = self.text_decoder(answer.input_ids, attention_mask=answer.attention_mask,
                                          encoder_hidden_states=question_output.last_hidden_state,
                                          encoder_attention_mask=quesiton.attention_mask.index_select(0, rows_of.long()),
                                          encoder_batch_index=rows_of, labels=answer_targets, return_dict=True, reduction="none",
                                          output_attentions=oa, output_hidden_states=oh, head_z=z["dec_head"], mlp_z=z["dec_mlp"])
        loss = ops.sum_scaled(weights * answer_output.loss, 1.0 / image.size(0))
        if not kd:
            return loss
        hidden_dict = {"image_hidden_states": image_hidden_states, "text_hidden_states": question_output.hidden_states,
                       "decoder_hidden_states": answer_output.hidden_states}
        attention_dict = {"image_attentions": image_attentions, "text_attentions": question_output.attentions,
                          "decoder_attentions": answer_output.attentions}
        cross_attention_dict = {"cross_attentions": question_output.cross_attentions,
                                "decoder_cross_attentions": answer_output.cross_attentions}
        return {"loss": loss, "hidden_dict": hidden_dict, "attention_dict": attention_dict, "cross_attention_dict": cross_attention_dict,
                "logits_dict": {"logits": answer_output.logits}}

    @torch.no_grad()
    def fake_forward(self, image, quesiton, answer=None, k=None, weights=None, train=False, output_attentions=None, output_hidden_states=None,
                     stop_prune=False):
        """model_generation.py:214-230: inference WITHOUT gates (the path for a model whose masks were materialised by
        utils/vqa_utils.py / efficientvlm_b200.prune).  The reference returns wall-clock seconds as the third item; timing is the
        caller's job here (CUDA events), so it is 0.0."""
        ids, probs = self.forward(image, quesiton, answer, k=k, train=False, use_gates=False)
        return ids, probs, 0.0

    @torch.no_grad()
    def rank_answer(self, question_states, question_atts, answer_ids, answer_atts, k, decoder_mask=None):
        """model_generation.py:233-300.  Returns (topk_ids [Q,k] indices into the answer list, topk_probs [Q,k], 0.0)."""
        decoder_head_z, decoder_mlp_z = decoder_mask if decoder_mask is not None else (None, None)
        num_ques = question_states.size(0)
        start_ids = answer_ids[0, 0].repeat(num_ques, 1)  # bos token
        start_output = self.text_decoder(start_ids, encoder_hidden_states=question_states, encoder_attention_mask=question_atts,
                                         return_dict=True, reduction="none", head_z=decoder_head_z, mlp_z=decoder_mlp_z)
        logits = start_output.logits[:, 0, :]  # first token's logit
        answer_first_token = answer_ids[:, 1]
        prob_first_token = torch.softmax(logits.float(), dim=1).index_select(dim=1, index=answer_first_token)
        topk_probs, topk_ids = prob_first_token.topk(k, dim=1)
        # answer input: [num_question * k, answer_len] (one gather instead of the per-question index_select loop)
        flat = topk_ids.reshape(-1)
        input_ids = answer_ids.index_select(0, flat)
        input_atts = answer_atts.index_select(0, flat)
        targets_ids = input_ids.masked_fill(input_ids == self.pad_token_id, -100)
        # candidate row r belongs to question r // k: indexed cross-attention instead of tile(question_states, 0, k)
        rows_of = torch.arange(num_ques, device=question_states.device, dtype=torch.int32).repeat_interleave(k)
        output = self.text_decoder(input_ids, attention_mask=input_atts, encoder_hidden_states=question_states,
                                   encoder_attention_mask=tile(question_atts, 0, k), encoder_batch_index=ops.UniformGroups(k, rows_of),
                                   labels=targets_ids,
                                   return_dict=True, reduction="none", head_z=decoder_head_z, mlp_z=decoder_mlp_z)
        answer_loss = output.loss.view(input_ids.size(0), -1)
        topk_probs = topk_probs.view(-1, 1)
        log_probs = torch.cat([topk_probs.log(), -answer_loss], dim=1)
        # re-calculate log probabilities for the answer sequences using chain rule
        log_probs_sum = log_probs.sum(1).view(num_ques, k)
        topk_probs = torch.softmax(log_probs_sum, dim=-1)
        topk_probs, rerank_id = topk_probs.topk(k, dim=1)
        topk_ids = torch.gather(topk_ids, 1, rerank_id)
        return topk_ids, topk_probs, 0.0


class EffXVLMForVQA(XVLMForVQA):
    """efficient_models/model_generation.py:23-300 — L0-gated VQA student (`l0_module` = VQAL0Module, eight gate types)."""
    gated = True


# ----------------------------------------------------------------------------------------------------------------------
# Eff_VQA.py:105-176 — KD terms and the loss mix of one pruning step
# ----------------------------------------------------------------------------------------------------------------------
def vqa_kd_losses(student_outputs, teacher_outputs, temperature=1.0):
    """All hidden / attention MSE terms of Eff_VQA.py:116-163 in ONE multi-pair launch + the logit KL (:164-167)."""
    sh, th = student_outputs["hidden_dict"], teacher_outputs["hidden_dict"]
    sa, ta = student_outputs["attention_dict"], teacher_outputs["attention_dict"]
    sc, tc = student_outputs["cross_attention_dict"], teacher_outputs["cross_attention_dict"]
    s_text_h = list(sh["text_hidden_states"])
    t_text_h = get_cor_teacher(th["text_hidden_states"], s_text_h)
    s_text_a = list(sa["text_attentions"])
    t_text_a = get_cor_teacher(ta["text_attentions"], s_text_a, is_attn=True)
    s_cross_a = list(sc["cross_attentions"])
    t_cross_a = get_cor_teacher(tc["cross_attentions"], s_cross_a, is_attn=True)
    s_img_h, s_img_a = list(sh["image_hidden_states"]), list(sa["image_attentions"])
    s_dec_h, s_dec_a, s_dec_c = list(sh["decoder_hidden_states"]), list(sa["decoder_attentions"]), list(sc["decoder_cross_attentions"])
    groups = [  # (name, students, mapped teachers, is_attn, is_img)   — slices 4 / 3 are literal in Eff_VQA.py:122,127,131-135
        ("text_hidden", s_text_h[:4], t_text_h[:4], False, False),
        ("text_attention", s_text_a[:3], t_text_a[:3], True, False),
        ("cross_hidden", s_text_h[4:], t_text_h[4:], False, False),
        ("cross_self_attention", s_text_a[3:], t_text_a[3:], True, False),
        ("cross_attention", s_cross_a, t_cross_a, True, False),
        ("image_hidden", s_img_h, get_cor_teacher(th["image_hidden_states"], s_img_h), False, True),
        ("image_attention", s_img_a, get_cor_teacher(ta["image_attentions"], s_img_a, is_attn=True), True, False),
        ("decoder_hidden", s_dec_h, get_cor_teacher(th["decoder_hidden_states"], s_dec_h), False, True),
        ("decoder_attention", s_dec_a, get_cor_teacher(ta["decoder_attentions"], s_dec_a, is_attn=True), True, False),
        ("decoder_cross", s_dec_c, get_cor_teacher(tc["decoder_cross_attentions"], s_dec_c, is_attn=True), True, False),
    ]
    S, T, W, spans = [], [], [], {}
    for name, s_list, t_list, is_attn, is_img in groups:
        s, t, w = _kd_pairs(s_list, t_list, is_attn, is_img)
        spans[name] = (len(S), len(S) + len(s))
        S += s
        T += t
        W += w
    per_pair = ops.mse_pairs(S, T, W)
    out = {name: per_pair[a:b].sum() for name, (a, b) in spans.items()}
    sl, tl = student_outputs["logits_dict"]["logits"], teacher_outputs["logits_dict"]["logits"]
    V = sl.shape[-1]
    s2, t2 = sl.reshape(-1, V), tl.reshape(-1, V)
    out["logits"] = ops.sum_scaled(ops.kl_rows(s2, t2, 1.0 / temperature), 1.0 / s2.shape[0])
    return out


def vqa_loss(student_outputs, teacher_outputs, l0_module=None, global_step=0, temperature=1.0):
    """`loss` of Eff_VQA.py:169-181: 0.4 * KD + 0.6 * task (+ Lagrangian).  Returns (loss, dict of logged components)."""
    kd = vqa_kd_losses(student_outputs, teacher_outputs, temperature)
    loss_small = student_outputs["loss"]
    loss_text_kd = kd["text_attention"] + kd["text_hidden"]
    loss_img_kd = kd["image_attention"] + kd["image_hidden"] * 0.2
    loss_cross_kd = (kd["cross_hidden"] + kd["cross_self_attention"] + kd["cross_attention"]) * 0.5
    loss_decoder_kd = kd["decoder_attention"] + kd["decoder_hidden"] + kd["decoder_cross"]
    loss_kd = kd["logits"] + loss_text_kd + loss_img_kd + loss_cross_kd + loss_decoder_kd
    loss = loss_kd * 0.4 + loss_small * 0.6
    parts = dict(loss_small=loss_small, loss_kd=loss_kd, loss_text_kd=loss_text_kd, loss_img_kd=loss_img_kd, loss_cross_kd=loss_cross_kd,
                 loss_decoder_kd=loss_decoder_kd, loss_logits_kd=kd["logits"], **{"kd_" + n: v for n, v in kd.items()})
    if l0_module is not None:
        lagrangian_loss, expected_sparsity, target_sparsity = l0_module.lagrangian_regularization(global_step)
        loss = loss + lagrangian_loss
        parts.update(loss_lagrangian=lagrangian_loss, expected_sparsity=expected_sparsity, target_sparsity=target_sparsity)
    return loss, parts


def set_vqa_teacher_attention_stride(teacher, student):
    """The KD terms read only every (teacher layers / student layers)-th teacher attention map (get_cor_teacher): tell the
    teacher's three encoders to materialise just those (extension; the skipped tuple entries are None)."""
    from .distill import set_teacher_attention_stride
    set_teacher_attention_stride(teacher, student)
    td, sd = teacher.text_decoder.bert.encoder, student.text_decoder.bert.encoder
    tn, sn = len(td.layer), len(sd.layer)
    td.attention_stride = tn // sn if sn > 0 and tn % sn == 0 and tn // sn > 1 else None
