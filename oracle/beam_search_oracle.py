"""ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-Python restatement of the beam search the reference's captioning evaluation runs:
`efficient_models/model_generation.py:474-483` calls `text_decoder.generate(num_beams=3, max_length=20, min_length=5, ...)`
(`Eff_Captioning.py:201-202`, `configs/x-vlm-small-ft/Captioning.yaml:29-31`), i.e. `GenerationMixin.generate -> beam_search` with
`BeamSearchScorer` / `BeamHypotheses` of **transformers 4.12.5** (`requirements.txt:2`).

PARITY UNPINNED: that package is not under /root/reference, the installed transformers 5.5 no longer ships these classes, and no test or
fixture of the reference pins a caption.  What follows restates the published 4.12.5 algorithm (generation_utils.py::beam_search,
generation_beam_search.py::{BeamSearchScorer.process, .finalize, BeamHypotheses.add, .is_done}, generation_logits_process.py::
{MinLengthLogitsProcessor, RepetitionPenaltyLogitsProcessor}) with its defaults at the reference's call site: length_penalty 1.0,
early_stopping False, one returned sequence, no n-gram / bad-word processors.  Python lists and floats only; `step_fn(ids)` is the model:
it maps the current [batch * beams, len] token lists to next-token LOG-probabilities (list of lists of floats).
"""
import math


class BeamHypotheses:
    def __init__(self, num_beams, length_penalty=1.0, early_stopping=False):
        self.num_beams, self.length_penalty, self.early_stopping = num_beams, length_penalty, early_stopping
        self.beams, self.worst_score = [], 1e9

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (len(hyp) ** self.length_penalty)
        if len(self.beams) < self.num_beams or score > self.worst_score:
            self.beams.append((score, list(hyp)))
            if len(self.beams) > self.num_beams:
                order = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[order[0][1]]
                self.worst_score = order[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self.beams) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


def beam_search(step_fn, input_ids, num_beams, max_length, min_length, pad_token_id, eos_token_id, vocab_size, repetition_penalty=1.0,
                length_penalty=1.0, early_stopping=False):
    """input_ids: list (batch) of token lists, already expanded `num_beams` times per item (generation_utils.py::_expand_inputs_for_generation).
    Returns the list (batch) of best sequences, padded like BeamSearchScorer.finalize."""
    batch = len(input_ids) // num_beams
    hyps = [BeamHypotheses(num_beams, length_penalty, early_stopping) for _ in range(batch)]
    done = [False] * batch
    beam_scores = [0.0 if b % num_beams == 0 else -1e9 for b in range(batch * num_beams)]
    ids = [list(r) for r in input_ids]
    last = None
    while True:
        cur_len = len(ids[0])
        logp = [list(r) for r in step_fn(ids)]
        for r, row in enumerate(logp):                                       # logits processors, in generate()'s order
            if repetition_penalty != 1.0:
                for tok in set(ids[r]):
                    row[tok] = row[tok] * repetition_penalty if row[tok] < 0 else row[tok] / repetition_penalty
            if cur_len < min_length:
                row[eos_token_id] = -math.inf
        next_scores, next_tokens, next_indices = [], [], []
        for b in range(batch):
            flat = [(logp[b * num_beams + k][v] + beam_scores[b * num_beams + k], k * vocab_size + v) for k in range(num_beams)
                    for v in range(vocab_size)]
            flat.sort(key=lambda t: (-t[0], t[1]))                           # torch.topk(largest, sorted); ties: lowest flat index first
            top = flat[:2 * num_beams]
            next_scores.append([s for s, _ in top])
            next_tokens.append([i % vocab_size for _, i in top])
            next_indices.append([i // vocab_size for _, i in top])
        new_scores, new_tokens, new_index = [], [], []
        for b in range(batch):                                               # BeamSearchScorer.process
            if done[b]:
                new_scores += [0.0] * num_beams
                new_tokens += [pad_token_id] * num_beams
                new_index += [0] * num_beams
                continue
            kept = 0
            for rank, (tok, sc, idx) in enumerate(zip(next_tokens[b], next_scores[b], next_indices[b])):
                bb = b * num_beams + idx
                if tok == eos_token_id:
                    if rank >= num_beams:
                        continue
                    hyps[b].add(ids[bb], sc)
                else:
                    new_scores.append(sc)
                    new_tokens.append(tok)
                    new_index.append(bb)
                    kept += 1
                if kept == num_beams:
                    break
            assert kept == num_beams
            done[b] = done[b] or hyps[b].is_done(max(next_scores[b]), cur_len)
        beam_scores = new_scores
        ids = [ids[i] + [t] for i, t in zip(new_index, new_tokens)]
        last = beam_scores
        if all(done) or len(ids[0]) >= max_length:
            break
    for b in range(batch):                                                   # BeamSearchScorer.finalize
        if done[b]:
            continue
        for k in range(num_beams):
            hyps[b].add(ids[b * num_beams + k], last[b * num_beams + k])
    best = [sorted(h.beams, key=lambda t: t[0])[-1][1] for h in hyps]
    lengths = [len(h) for h in best]
    out_len = min(max(lengths) + 1, max_length)
    out = []
    for h in best:
        row = [pad_token_id] * out_len
        row[:len(h)] = h[:out_len]
        if len(h) < max_length:
            row[len(h)] = eos_token_id
        out.append(row)
    return out
