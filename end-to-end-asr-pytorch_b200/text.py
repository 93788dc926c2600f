"""Token <-> text codecs with the reference's conventions (/root/reference/src/text.py:33-42,59): <pad>=0 (also the
CTC blank), <eos>=1, <unk>=2; `encode` appends <eos>; `decode(ids, ignore_repeat)` drops pads, collapses repeats when
asked (CTC greedy path) and stops at <eos>.  CPU string work - not on the accelerated path."""


class _Codec:
    pad_idx, eos_idx, unk_idx = 0, 1, 2
    token_type = "?"

    def _crop(self, ids, ignore_repeat):
        out = []
        for t, i in enumerate(ids):
            if i == self.eos_idx:
                break
            if i == self.pad_idx or (ignore_repeat and t > 0 and i == ids[t - 1]):
                continue
            out.append(i)
        return out

    def __repr__(self):
        return "<{} vocab_size={}>".format(type(self).__name__, self.vocab_size)


class CharacterTextEncoder(_Codec):
    token_type = "character"

    def __init__(self, vocab_list):
        self._vocab = ["<pad>", "<eos>", "<unk>"] + list(vocab_list)
        self._index = {v: i for i, v in enumerate(self._vocab)}

    @classmethod
    def load_from_file(cls, vocab_file):
        with open(vocab_file, "r") as f:
            return cls([line.strip("\r\n") for line in f])     # keep the space token

    @property
    def vocab_size(self):
        return len(self._vocab)

    def encode(self, s):
        return [self._index.get(ch, self.unk_idx) for ch in s.strip("\r\n ")] + [self.eos_idx]

    def decode(self, ids, ignore_repeat=False):
        # note: the reference's character decoder tests <pad>/repeat before <eos>; same result on valid input
        return "".join(self._vocab[i] for i in self._crop(ids, ignore_repeat))


class WordTextEncoder(CharacterTextEncoder):
    token_type = "word"

    def encode(self, s):
        return [self._index.get(w, self.unk_idx) for w in s.strip("\r\n ").split(" ")] + [self.eos_idx]

    def decode(self, ids, ignore_repeat=False):
        return " ".join(self._vocab[i] for i in self._crop(ids, ignore_repeat))


class SubwordTextEncoder(_Codec):
    token_type = "subword"

    def __init__(self, spm):
        if spm.pad_id() != 0 or spm.eos_id() != 1 or spm.unk_id() != 2:
            raise ValueError("sentencepiece model must be trained with --pad_id=0 --eos_id=1 --unk_id=2 --bos_id=-1")
        self.spm = spm

    @classmethod
    def load_from_file(cls, filepath):
        import sentencepiece as splib
        spm = splib.SentencePieceProcessor()
        spm.load(filepath)
        spm.set_encode_extra_options(":eos")
        return cls(spm)

    @property
    def vocab_size(self):
        return len(self.spm)

    def encode(self, s):
        return self.spm.encode_as_ids(s)

    def decode(self, ids, ignore_repeat=False):
        return self.spm.decode_ids(self._crop(ids, ignore_repeat))


def load_text_encoder(mode, vocab_file):
    if mode == "character":
        return CharacterTextEncoder.load_from_file(vocab_file)
    if mode == "subword":
        return SubwordTextEncoder.load_from_file(vocab_file)
    if mode == "word":
        return WordTextEncoder.load_from_file(vocab_file)
    raise NotImplementedError("Unsupported text encoder mode `{}` (bert is outside the hot path)".format(mode))
