"""CPU tests of slamkit_b200/generation.py against `transformers`' own generation code: the logits processing equals HF's
warpers, and greedy decoding of a LEFT-padded batch (the way SpeechLM.generate calls it, slamkit/model/speech_lm.py:38-55)
equals `Qwen2ForCausalLM.generate` on the same random model, including the eos / pad tail and `bad_words_ids`."""
import pytest
import torch

from slamkit_b200.generation import generate_tokens, process_logits, select_next


def test_logits_processing_equals_hf_warpers():
    from transformers.generation.logits_process import (NoBadWordsLogitsProcessor, TemperatureLogitsWarper, TopKLogitsWarper,
                                                       TopPLogitsWarper)
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(4, 502, generator=g) * 3
    ids = torch.zeros(4, 1, dtype=torch.long)
    for temp, k, p, bad in [(0.8, 25, None, None), (1.0, None, 0.9, [3, 7]), (0.5, 10, 0.7, [0]), (1.3, 600, 0.95, None)]:
        want = logits.clone()
        if bad:
            want = NoBadWordsLogitsProcessor([[b] for b in bad], eos_token_id=None)(ids, want)
        if temp != 1.0:
            want = TemperatureLogitsWarper(temp)(ids, want)
        if k:
            want = TopKLogitsWarper(k)(ids, want)
        if p:
            want = TopPLogitsWarper(p)(ids, want)
        got = process_logits(logits, temp, k, p, bad)
        assert torch.equal(torch.isinf(got), torch.isinf(want))
        keep = ~torch.isinf(want)
        assert torch.allclose(got[keep], want[keep], rtol=0, atol=1e-6)
    # sampling draws from exactly that distribution
    s = process_logits(logits[0], 0.8, 5, None, None)
    allowed = set(torch.nonzero(~torch.isinf(s)).flatten().tolist())
    gen = torch.Generator().manual_seed(1)
    assert all(select_next(logits[0], True, 0.8, 5, None, None, gen) in allowed for _ in range(50)) and len(allowed) == 5


@pytest.fixture(scope="module")
def tiny_hf():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(3)
    cfg = Qwen2Config(vocab_size=502, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=1, max_position_embeddings=256, tie_word_embeddings=True, pad_token_id=0,
                      bos_token_id=1, eos_token_id=1)
    m = Qwen2ForCausalLM(cfg).eval()
    return m


def _left_padded_batch():
    g = torch.Generator().manual_seed(5)
    a = torch.randint(2, 502, (9,), generator=g)
    b = torch.randint(2, 502, (5,), generator=g)
    ids = torch.zeros(2, 9, dtype=torch.long)
    mask = torch.zeros(2, 9, dtype=torch.long)
    ids[0], mask[0] = a, 1
    ids[1, 4:], mask[1, 4:] = b, 1
    return ids, mask


def test_greedy_left_padded_batch_equals_hf_generate(tiny_hf):
    ids, mask = _left_padded_batch()

    def next_logits(x):
        with torch.no_grad():
            return tiny_hf(input_ids=x).logits[0, -1]

    with torch.no_grad():
        want = tiny_hf.generate(input_ids=ids, attention_mask=mask, do_sample=False, max_new_tokens=8, eos_token_id=None,
                                pad_token_id=0)
    got = generate_tokens(next_logits, ids, attention_mask=mask, do_sample=False, max_new_tokens=8, pad_token_id=0)
    assert torch.equal(got, want), (got, want)
    # eos: stop row 0 at its 3rd new token, keep decoding row 1, pad the tail; plus a banned token
    eos = int(want[0, 9 + 2])
    ban = int(want[1, 9])                    # row 1's first greedy choice is banned -> both implementations must move on
    with torch.no_grad():
        want2 = tiny_hf.generate(input_ids=ids, attention_mask=mask, do_sample=False, max_new_tokens=8, eos_token_id=eos,
                                 pad_token_id=0, bad_words_ids=[[ban]])
    got2 = generate_tokens(next_logits, ids, attention_mask=mask, do_sample=False, max_new_tokens=8, eos_token_id=eos,
                           pad_token_id=0, bad_words_ids=[[ban]])
    assert torch.equal(got2, want2), (got2, want2)
    assert ban not in got2[:, 9:].tolist()[1]
    # max_length counts the prompt
    got3 = generate_tokens(next_logits, ids, attention_mask=mask, max_length=12, pad_token_id=0)
    assert got3.shape == (2, 12) and torch.equal(got3, want[:, :12])


def test_argument_errors():
    f = lambda x: torch.zeros(502)
    ids = torch.ones(1, 4, dtype=torch.long)
    with pytest.raises(ValueError, match="left-padding"):
        generate_tokens(f, ids, attention_mask=torch.tensor([[1, 1, 0, 0]]), max_new_tokens=2)
    with pytest.raises(NotImplementedError, match="single-token"):
        generate_tokens(f, ids, max_new_tokens=2, bad_words_ids=[[3, 4]])
    with pytest.raises(ValueError, match="longer than max_length"):
        generate_tokens(f, ids, max_length=2)
    out = generate_tokens(f, ids, max_new_tokens=5, max_positions=6)       # the RoPE table bounds the total length
    assert out.shape == (1, 6)
