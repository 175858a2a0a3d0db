"""Character tokenizer of the sampling path.

Same mapping as the reference `Tokenizer` (diffusion_handwriting_generation/
tokenizer.py:8-34): alphabet "_" + a-z + A-Z + 0-9 + ".?!,'\"- " -> ids 2..72 in
that order, id 0 = padding, id 1 = end of sentence (appended by `encode`),
unknown characters -> 2 ("_").
"""
import string

ALPHABET = "_" + string.ascii_letters + string.digits + ".?!,'\"- "
PAD_ID, END_ID, UNKNOWN_ID = 0, 1, 2
VOCAB_SIZE = len(ALPHABET) + 2  # 73, the embedding table height (text_style.py:70)


class Tokenizer:
    def __init__(self):
        self.text = ALPHABET
        self.vocab_size = VOCAB_SIZE
        self.tokens = {ch: i + 2 for i, ch in enumerate(ALPHABET)}
        self.chars = {i + 2: ch for i, ch in enumerate(ALPHABET)}
        self.chars[PAD_ID], self.chars[END_ID] = " ", "<end>"

    def encode(self, text):
        ids = [self.tokens.get(ch, UNKNOWN_ID) for ch in text]
        ids.append(END_ID)
        return ids

    def decode(self, tokens):
        if hasattr(tokens, "tolist"):
            tokens = tokens.tolist()
        return "".join(self.chars[int(t)] for t in tokens)


def stroke_length(n_tokens):
    """T = 16 * n_tokens rounded up to the next multiple of 8 strictly above
    (inference.py:77-78: T = T - T % 8 + 8)."""
    t = 16 * n_tokens
    return t - (t % 8) + 8
