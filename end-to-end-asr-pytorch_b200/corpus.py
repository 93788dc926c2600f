"""LibriSpeech file listing (/root/reference/corpus/librispeech.py:28-61): glob the audio files of the splits, read
`<spk>-<chap>.trans.txt`, sort by token length, serve single items or buckets.  Out of the accelerated path; kept so
that `main.py --config <libri yaml>` is a drop-in."""
from pathlib import Path

from torch.utils.data import Dataset


def _transcripts(directory):
    table = {}
    for f in Path(directory).glob("*.trans.txt"):
        for line in open(f, "r"):
            utt, _, txt = line.rstrip("\n").partition(" ")
            table[utt] = txt
    return table


class LibriDataset(Dataset):
    def __init__(self, path, split, tokenizer, bucket_size, ascending=False):
        self.path, self.bucket_size = path, bucket_size
        files = []
        for s in split:
            found = sorted(list(Path(path, s).rglob("*.flac")) + list(Path(path, s).rglob("*.wav")))
            assert len(found) > 0, "No data found @ {}".format(Path(path, s))
            files += found
        cache = {}
        texts = []
        for f in files:
            d = str(f.parent)
            if d not in cache:
                cache[d] = _transcripts(d)
            texts.append(tokenizer.encode(cache[d][f.name.split(".")[0]]))
        pairs = sorted(zip(files, texts), reverse=not ascending, key=lambda x: len(x[1]))
        self.file_list, self.text = [p[0] for p in pairs], [p[1] for p in pairs]

    def __getitem__(self, index):
        if self.bucket_size > 1:
            index = min(len(self.file_list) - self.bucket_size, index)
            return list(zip(self.file_list[index:index + self.bucket_size], self.text[index:index + self.bucket_size]))
        return self.file_list[index], self.text[index]

    def __len__(self):
        return len(self.file_list)
