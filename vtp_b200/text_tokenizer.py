"""Caption tokenisation for the contrastive objective: the byte-level BPE of
`vtp/tokenizers/text_tokenizer.py` (class `SimpleTokenizer`, lines 144-295 — OpenAI CLIP's tokenizer) with the same
surface (`encode`, `decode`, `__call__`, `vocab_size`, `sot_token_id`, `eot_token_id`, `all_special_ids`,
`context_length`) and the same token ids on every input, built for the step's cadence: 256 captions have to be ready
every ~200 ms beside the GPU work (`vtp_b200/data.py` runs this in a worker thread under the previous step).

What is different underneath:
  * merging runs on INTEGER symbol ids with a `(left id, right id) -> (rank, merged id)` table: no string
    concatenation, no `' '.join / split`, no per-merge tuple rebuilds of strings;
  * a word's id sequence is cached per surface form (captions repeat words), whole captions are cached too (datasets
    repeat captions across epochs) with a bounded size;
  * a batch goes straight into one preallocated int64 array (start token, ids, end token, zero padding; an over-long
    caption is cut to the context length and its last slot becomes the end token, text_tokenizer.py:286-292).

The vocabulary file (`bpe_simple_vocab_16e6.txt.gz`, shipped by the reference under `tools/`) is NOT part of this
repository: pass its path, or set `VTP_BPE_PATH`, or have it next to the caller — the lookup mirrors
text_tokenizer.py:37-72."""
from __future__ import annotations

import gzip
import html
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

try:  # same optional dependencies as the reference (text_tokenizer.py:13-26): behaviour must not differ where they exist
    import ftfy

    def _fix_text(t: str) -> str:
        return ftfy.fix_text(t)
except ImportError:  # pragma: no cover - depends on the environment
    def _fix_text(t: str) -> str:
        return t

try:
    import regex as _re

    _UNICODE_CLASSES = True
except ImportError:  # pragma: no cover
    import re as _re

    _UNICODE_CLASSES = False

DEFAULT_CONTEXT_LENGTH = 77
_N_MERGES = 49152 - 256 - 2   # merges kept from the vocabulary file (text_tokenizer.py:173)
_WORD_END = "</w>"


def _byte_alphabet() -> List[str]:
    """The 256 printable stand-ins of the byte values, in the order the vocabulary numbers them (ids 0..255): the bytes
    that are printable Latin-1 keep their own character and come first, in byte order; every other byte gets the next
    code point from 256 upwards (text_tokenizer.py:76-95)."""
    keep = [b for b in range(256) if 33 <= b <= 126 or 161 <= b <= 172 or 174 <= b <= 255]
    kept = set(keep)
    chars = [chr(b) for b in keep]
    chars += [chr(256 + i) for i, _ in enumerate(b for b in range(256) if b not in kept)]
    return chars


def _byte_to_id() -> List[int]:
    """byte value -> vocabulary id of its stand-in character."""
    keep = [b for b in range(256) if 33 <= b <= 126 or 161 <= b <= 172 or 174 <= b <= 255]
    kept = set(keep)
    order = keep + [b for b in range(256) if b not in kept]
    table = [0] * 256
    for idx, b in enumerate(order):
        table[b] = idx
    return table


def find_bpe_file(explicit: Optional[str] = None) -> str:
    """`explicit`, then $VTP_BPE_PATH, then the places the reference looks in relative to ITS package (a checkout on
    sys.path), then the working directory."""
    cands: List[str] = []
    if explicit:
        cands.append(explicit)
    if os.environ.get("VTP_BPE_PATH"):
        cands.append(os.environ["VTP_BPE_PATH"])
    name = "bpe_simple_vocab_16e6.txt.gz"
    try:  # a reference checkout on the path (compat shim or plain): <checkout>/tools/<name> and next to its tokenizer
        import importlib.util

        spec = importlib.util.find_spec("vtp.tokenizers")
        for loc in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
            root = os.path.dirname(os.path.dirname(loc))
            cands += [os.path.join(loc, name), os.path.join(os.path.dirname(loc), name), os.path.join(root, name),
                      os.path.join(root, "tools", name)]
    except (ImportError, ValueError, AttributeError):
        pass
    cands += [os.path.join(os.getcwd(), name), os.path.join(os.getcwd(), "tools", name)]
    for c in cands:
        if c and os.path.exists(c):
            return os.path.abspath(c)
    raise FileNotFoundError(
        "BPE vocabulary file not found. Pass bpe_path=, set VTP_BPE_PATH, or put bpe_simple_vocab_16e6.txt.gz where "
        "the reference keeps it (its tools/ directory).")


def _clean(text: str, lower: bool) -> str:
    """text_tokenizer.py:110-131: ftfy (when installed), two rounds of HTML unescaping, strip, collapse whitespace."""
    text = html.unescape(html.unescape(_fix_text(text))).strip()
    text = " ".join(text.split())
    return text.lower() if lower else text


class BPETokenizer:
    """Byte-level BPE with the reference `SimpleTokenizer`'s vocabulary layout: ids 0..255 the byte stand-ins, 256..511
    the same with the end-of-word marker, then one id per merge in file order, then the special tokens."""

    def __init__(self, bpe_path: Optional[str] = None, additional_special_tokens: Optional[List[str]] = None,
                 context_length: Optional[int] = DEFAULT_CONTEXT_LENGTH, clean: str = "lower",
                 caption_cache_size: int = 1 << 16):
        if bpe_path is None:
            path = find_bpe_file()
        elif not os.path.exists(bpe_path):
            raise FileNotFoundError(f"BPE vocabulary file not found at {bpe_path}. Please ensure the file exists.")
        else:
            path = bpe_path
        with gzip.open(path) as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(l.split()) for l in lines[1:_N_MERGES + 1]]
        alphabet = _byte_alphabet()
        symbols: List[str] = alphabet + [c + _WORD_END for c in alphabet] + ["".join(m) for m in merges]
        self.special_tokens: List[str] = ["<start_of_text>", "<end_of_text>"] + list(additional_special_tokens or [])
        symbols += self.special_tokens
        # like the reference's dict(zip(vocab, range(len(vocab)))): a repeated string keeps its LAST id
        self.encoder: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
        self.decoder: Dict[int, str] = {i: s for s, i in self.encoder.items()}
        self._symbols = symbols
        # (left id, right id) -> (rank, merged id).  Ids of the operands are looked up by STRING so that they agree with the
        # encoder the reference builds (first occurrence wins for a duplicated merge, as in dict(zip(merges, ranks)) ...
        # where the LAST rank wins: keep that rule too)
        pair_table: Dict[Tuple[int, int], Tuple[int, int]] = {}
        for rank, m in enumerate(merges):
            if len(m) != 2:
                continue
            a, b = self.encoder.get(m[0]), self.encoder.get(m[1])
            if a is None or b is None:
                continue
            pair_table[(a, b)] = (rank, self.encoder[m[0] + m[1]])
        self._pairs = pair_table
        self._byte_id = _byte_to_id()
        self._word_cache: Dict[str, Tuple[int, ...]] = {t: (self.encoder[t],) for t in self.special_tokens}
        self._caption_cache: Dict[str, Tuple[int, ...]] = {}
        self._caption_cache_size = caption_cache_size
        special = "|".join(self.special_tokens)   # (unescaped, as in the reference: special tokens are plain words)
        if _UNICODE_CLASSES:
            body = r"""|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""
        else:  # pragma: no cover
            body = r"""|'s|'t|'re|'ve|'m|'ll|'d|[a-zA-Z]+|[0-9]+|[^\s\w]+"""
        self.pat = _re.compile(special + body, _re.IGNORECASE)
        self.vocab_size = len(self.encoder)
        self.all_special_ids = [self.encoder[t] for t in self.special_tokens]
        self.sot_token_id, self.eot_token_id = self.all_special_ids[0], self.all_special_ids[1]
        self.context_length = context_length
        self._lower = clean != "whitespace"     # 'lower' and any unknown value lower-case (text_tokenizer.py:134-141)
        self.byte_decoder: Dict[str, int] = {alphabet[self._byte_id[b]]: b for b in range(256)}

    # ------------------------------------------------------------------------------------------------ one word
    def _merge_word(self, word: str) -> Tuple[int, ...]:
        """Greedy lowest-rank merging (text_tokenizer.py:208-248) on id sequences.  Per round the lowest-ranked adjacent
        pair present is merged at EVERY non-overlapping occurrence, left to right."""
        hit = self._word_cache.get(word)
        if hit is not None:
            return hit
        bid = self._byte_id
        ids = [bid[b] for b in word.encode("utf-8")]
        ids[-1] += 256                                   # end-of-word variant of the last symbol
        pairs = self._pairs
        while len(ids) > 1:
            best_rank, best = None, None
            for k in range(len(ids) - 1):
                e = pairs.get((ids[k], ids[k + 1]))
                if e is not None and (best_rank is None or e[0] < best_rank):
                    best_rank, best = e[0], (ids[k], ids[k + 1], e[1])
            if best is None:
                break
            a, b, merged = best
            out: List[int] = []
            k, n = 0, len(ids)
            while k < n:
                if k + 1 < n and ids[k] == a and ids[k + 1] == b:
                    out.append(merged)
                    k += 2
                else:
                    out.append(ids[k])
                    k += 1
            ids = out
        res = tuple(ids)
        self._word_cache[word] = res
        return res

    # ------------------------------------------------------------------------------------------------ API
    def encode(self, text: str) -> List[int]:
        """Token ids of one caption, without the start / end tokens (text_tokenizer.py:250-257)."""
        hit = self._caption_cache.get(text)
        if hit is not None:
            return list(hit)
        out: List[int] = []
        for word in self.pat.findall(_clean(text, self._lower)):
            out.extend(self._merge_word(word))
        if len(self._caption_cache) < self._caption_cache_size:
            self._caption_cache[text] = tuple(out)
        return out

    def decode(self, tokens: Sequence[int]) -> str:
        """text_tokenizer.py:259-263."""
        text = "".join(self.decoder[int(t)] for t in tokens)
        # the end-of-word marker (and the special tokens) are plain ASCII and survive the byte mapping unchanged: map every
        # character back to its byte, decode, then turn the markers into spaces
        data = bytearray(self.byte_decoder[c] for c in text)
        return data.decode("utf-8", errors="replace").replace(_WORD_END, " ")

    def __call__(self, texts: Union[str, Sequence[str]], context_length: Optional[int] = None) -> torch.LongTensor:
        """int64 [len(texts), context_length]: <start_of_text> ids <end_of_text>, zero padded; over-long captions are cut
        and end with <end_of_text> (text_tokenizer.py:265-294)."""
        if isinstance(texts, str):
            texts = [texts]
        L = context_length or self.context_length
        assert L, "Please set a valid context length"
        out = np.zeros((len(texts), L), dtype=np.int64)
        sot, eot = self.sot_token_id, self.eot_token_id
        for i, t in enumerate(texts):
            ids = self.encode(t)
            n = len(ids) + 2
            if n > L:
                row = [sot] + ids[:L - 1]
                row[L - 1] = eot
                out[i, :] = row[:L]
            else:
                out[i, 0] = sot
                if ids:
                    out[i, 1:n - 1] = ids
                out[i, n - 1] = eot
        return torch.from_numpy(out)


# names the reference exports (text_tokenizer.py:144,297)
SimpleTokenizer = BPETokenizer


def get_tokenizer(model_name: str = "ViT-B-32", context_length: Optional[int] = None, cache_dir: Optional[str] = None,
                  **kwargs) -> BPETokenizer:
    """text_tokenizer.py:297-325 (`model_name` / `cache_dir` are accepted for compatibility and unused, as upstream)."""
    return BPETokenizer(context_length=context_length or DEFAULT_CONTEXT_LENGTH, **kwargs)
