"""TEST INFRASTRUCTURE ONLY — a tiny whitespace tokenizer with the BertTokenizer surface the captioning models use
(bert-base-uncased's vocabulary files are not available offline).  Word ids are a CRC of the word, so the reference-side fixture
generator and the tests agree without a vocabulary file."""
import zlib

import torch


class Encoding:
    def __init__(self, input_ids, attention_mask):
        self.input_ids, self.attention_mask = input_ids, attention_mask

    def to(self, device):
        if torch.is_tensor(self.input_ids):
            return Encoding(self.input_ids.to(device), self.attention_mask.to(device))
        return self


class FakeTokenizer:
    cls_token, sep_token, pad_token = "[CLS]", "[SEP]", "[PAD]"
    pad_token_id, cls_token_id, sep_token_id = 0, 2, 3

    def __init__(self, vocab_size=211):
        self.vocab_size = vocab_size
        self.words = {}

    def add_special_tokens(self, mapping):
        self.bos_token, self.eos_token = mapping.get("bos_token"), mapping.get("eos_token")
        self.bos_token_id, self.eos_token_id = self.cls_token_id, self.sep_token_id

    def _word_id(self, w):
        i = 4 + zlib.crc32(w.encode()) % (self.vocab_size - 4)
        self.words.setdefault(i, w)
        return i

    def _encode(self, text, max_length=None):
        ids = [self.cls_token_id] + [self._word_id(w) for w in text.lower().split()] + [self.sep_token_id]
        if max_length is not None and len(ids) > max_length:
            ids = ids[:max_length - 1] + [self.sep_token_id]
        return ids

    def __call__(self, text, padding=None, truncation=False, max_length=None, return_tensors=None):
        if isinstance(text, str):
            ids = self._encode(text, max_length if truncation else None)
            if return_tensors == "pt":
                return Encoding(torch.tensor([ids]), torch.ones(1, len(ids), dtype=torch.long))
            return Encoding(ids, [1] * len(ids))
        rows = [self._encode(t, max_length if truncation else None) for t in text]
        n = max(len(r) for r in rows)
        input_ids = torch.tensor([r + [self.pad_token_id] * (n - len(r)) for r in rows])
        attention_mask = torch.tensor([[1] * len(r) + [0] * (n - len(r)) for r in rows])
        return Encoding(input_ids, attention_mask)

    def decode(self, ids, skip_special_tokens=True):
        out = []
        for i in (ids.tolist() if torch.is_tensor(ids) else ids):
            if skip_special_tokens and i in (self.pad_token_id, self.cls_token_id, self.sep_token_id):
                continue
            out.append(self.words.get(i, "w%d" % i))
        return " ".join(out)
