"""Caption tokenizer (vtp_b200/text_tokenizer.py) — CPU tests.

  * against the LIVE reference `SimpleTokenizer` (vtp/tokenizers/text_tokenizer.py:144-295) with the reference's own
    vocabulary file: identical vocabulary, identical ids for every caption of a corpus built to hit the pattern's
    branches, identical truncation and decoding.  Runs only where /root/reference exists (the dev container); the
    vocabulary is the reference's data file and is not committed;
  * self-contained cases on a tiny synthetic vocabulary (written to a temp dir): merge order, end-of-word variants,
    special tokens, truncation rule, errors."""
import gzip
import importlib.util
import os
import random

import pytest
import torch

from vtp_b200.text_tokenizer import BPETokenizer, find_bpe_file, get_tokenizer

REF_TOK = "/root/reference/vtp/tokenizers/text_tokenizer.py"
REF_BPE = "/root/reference/tools/bpe_simple_vocab_16e6.txt.gz"

CORPUS = [
    "a photo of a cat", "A Photo of a CAT!!!", "  multiple   spaces\tand\nnewlines ",
    "it's the dog's ball, they've won; I'm here, we'll go, he'd say, don't", "'s't're've'm'll'd", "''''",
    "naïve café déjà vu — “quotes” ‘single’ … ellipsis", "日本語のテキスト と 中文文本 and한국어", "emoji 😀😃 🤖👍🏽 flags 🇩🇪",
    "numbers 1234567890 3.14159 1e-5 ½ ²", "&amp;lt;b&amp;gt; html &amp; entities &lt;i&gt; &#39;x&#39;",
    "<start_of_text> literal specials <end_of_text> inside", "<START_OF_TEXT> upper special", "", "   ", "x", "a" * 300,
    "word " * 200, "supercalifragilisticexpialidocious antidisestablishmentarianism",
    "e-mail: someone@example.com, http://example.com/path?query=1&b=2", "tabs\tand\x00control\x07chars",
    "mixed123abc456 under_score-dash", "ÀÉÎÕÜ ßẞ ǅ İi", "𝔘𝔫𝔦𝔠𝔬𝔡𝔢 math 𝟘𝟙𝟚", "！？。、", "á combining ë",
]


def _load_reference_tokenizer():
    spec = importlib.util.spec_from_file_location("_ref_text_tokenizer", REF_TOK)   # by path: `import vtp` needs omegaconf
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not (os.path.exists(REF_TOK) and os.path.exists(REF_BPE)), reason="needs the reference checkout")
def test_token_ids_equal_the_live_reference():
    ref_mod = _load_reference_tokenizer()
    ours, ref = BPETokenizer(REF_BPE), ref_mod.SimpleTokenizer(REF_BPE)
    assert ours.encoder == ref.encoder and ours.decoder == ref.decoder and ours.byte_decoder == ref.byte_decoder
    assert (ours.vocab_size, ours.sot_token_id, ours.eot_token_id, ours.all_special_ids, ours.context_length) == \
           (ref.vocab_size, ref.sot_token_id, ref.eot_token_id, ref.all_special_ids, ref.context_length)
    rng = random.Random(0)
    alphabet = "abcdefghijklmnopqrstuvwxyz  ABC.,!?'0123456789-éüñ日本😀"
    corpus = CORPUS + ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 120))) for _ in range(400)]
    for text in corpus:
        a, b = ours.encode(text), ref.encode(text)
        assert a == b, text
        assert ours.decode(a) == ref.decode(b)
    for L in (None, 8, 16, 77, 200):            # padding, exact fit, truncation (last slot becomes <end_of_text>)
        x, y = ours(corpus, L), ref(corpus, L)
        assert x.dtype == torch.long and torch.equal(x, y)
    assert torch.equal(ours("one caption"), ref("one caption"))
    # second call: served from the caption cache, same ids
    assert torch.equal(ours(corpus), ref(corpus))
    # no lower-casing + an extra special token (case-sensitive cache hit of specials, as upstream)
    o2 = BPETokenizer(REF_BPE, clean="whitespace", additional_special_tokens=["<mask>"])
    r2 = ref_mod.SimpleTokenizer(REF_BPE, clean="whitespace", additional_special_tokens=["<mask>"])
    extra = corpus + ["Keep CASE <mask> <Mask> <start_of_text>"]
    assert torch.equal(o2(extra), r2(extra)) and o2.vocab_size == r2.vocab_size == ours.vocab_size + 1
    # the lookup finds the reference's copy when a checkout is importable, and get_tokenizer mirrors the factory
    assert get_tokenizer(bpe_path=REF_BPE, context_length=32)("a cat").shape == (1, 32)


def _tiny_vocab(tmp_path):
    """header line + five merges: 't h', 'th e</w>', 'c a', 'ca t</w>', 'a t</w>' (the last can never apply after 'c a')."""
    p = tmp_path / "tiny_bpe.txt.gz"
    with gzip.open(p, "wb") as f:
        f.write('"version"\nt h\nth e</w>\nc a\nca t</w>\na t</w>\n'.encode("utf-8"))
    return str(p)


def test_merge_order_and_layout_on_a_synthetic_vocabulary(tmp_path):
    tok = BPETokenizer(_tiny_vocab(tmp_path), context_length=8)
    # layout: 256 byte symbols, 256 end-of-word variants, merges in file order (a trailing empty line yields one more,
    # empty, entry exactly as upstream's split('\n')), then the two special tokens
    assert tok.encoder["th"] == 512 and tok.encoder["the</w>"] == 513 and tok.encoder["cat</w>"] == 515
    assert tok.sot_token_id == tok.vocab_size - 2 and tok.eot_token_id == tok.vocab_size - 1
    b = lambda ch: tok.encoder[ch]
    assert tok.encode("the") == [513]                                   # t h -> th ; th e</w> -> the</w>
    assert tok.encode("The  cat") == [513, 515]                         # lower-cased, whitespace collapsed
    assert tok.encode("that") == [512, tok.encoder["at</w>"]]           # 't h' (rank 0) before 'a t</w>' (rank 4)
    assert tok.encode("tht") == [512, b("t</w>")]                       # no merge for (th, t</w>)
    assert tok.encode("ththe") == [512, 513]                            # every occurrence of the best pair merges per round
    assert tok.encode("é") == [b("Ã"), b("©") + 256]                    # two UTF-8 bytes, the last one end-of-word
    assert tok.encode("<end_of_text>") == [tok.eot_token_id]
    assert tok.decode(tok.encode("the cat é")) == "the cat é "
    out = tok(["the", "the cat the cat the cat the cat", ""])
    sot, eot = tok.sot_token_id, tok.eot_token_id
    assert out.tolist() == [[sot, 513, eot, 0, 0, 0, 0, 0],
                            [sot, 513, 515, 513, 515, 513, 515, eot],    # cut to 8, last slot = end token
                            [sot, eot, 0, 0, 0, 0, 0, 0]]
    assert tok(["the"], context_length=2).tolist() == [[sot, eot]]


def test_vocabulary_lookup_and_errors(tmp_path, monkeypatch):
    with pytest.raises(FileNotFoundError):
        BPETokenizer(str(tmp_path / "missing.txt.gz"))
    monkeypatch.setenv("VTP_BPE_PATH", _tiny_vocab(tmp_path))
    assert find_bpe_file() == os.path.abspath(_tiny_vocab(tmp_path))
    assert BPETokenizer().encode("the") == [513]
    with pytest.raises(AssertionError):
        BPETokenizer(context_length=None)("x")
