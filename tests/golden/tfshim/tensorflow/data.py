"""tf.data subset: an eager, in-order stand-in for the pipeline of sagan/dataset.py:16,36-38 (take, shuffle, map,
batch with drop_remainder).  `shuffle` keeps the order (the fixture wants reproducible batches; shuffling is not part
of the per-record arithmetic)."""
import numpy as np

from ._core import Tensor, raw

records = {}            # file name -> list of stand-in records, registered by the fixture script


class Dataset:
    def __init__(self, items):
        self.items = list(items)

    def take(self, n):
        return Dataset(self.items if n is None or int(n) < 0 else self.items[:int(n)])

    def shuffle(self, buffer_size, *a, **k):
        return Dataset(self.items)

    def map(self, fn, *a, **k):
        return Dataset([fn(it) for it in self.items])

    def batch(self, n, drop_remainder=False):
        out = []
        for i in range(0, len(self.items), n):
            chunk = self.items[i:i + n]
            if len(chunk) < n and drop_remainder:
                break
            cols = list(zip(*chunk))
            out.append(tuple(Tensor(np.stack([raw(c) for c in col]), keep_dtype=True) for col in cols))
        return Dataset(out)

    def __iter__(self):
        return iter(self.items)


class TFRecordDataset(Dataset):
    def __init__(self, filenames):
        names = [filenames] if isinstance(filenames, str) else list(filenames)
        items = []
        for f in sorted(names):
            items += records[f]
        super().__init__(items)
