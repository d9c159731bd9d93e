"""-m "not gpu": minibatch index semantics (bit-exact contract).

spvipes_b200.trainer.split_indices / epoch_batches restate reference data/_multi_datasplitter.py:65-85,
dataloaders/_ann_dataloader.py:85-92 and dataloaders/_concat_dataloader.py:108-110.  Here they are compared, index for
index, with the REAL torch DataLoader(BatchSampler(RandomSampler)) / zip / itertools.cycle stack the reference builds, run
under the same torch seed, and with numpy's RandomState permutation for the split."""
import math
from itertools import cycle

import numpy as np
import pytest
import torch
from torch.utils.data import BatchSampler, DataLoader, Dataset, RandomSampler, SequentialSampler


class _Rows(Dataset):
    """like scvi's AnnTorchDataset restricted to `indices`: __getitem__ takes the whole list of positions of a minibatch"""

    def __init__(self, indices):
        self.indices = np.asarray(indices)

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, pos):
        return self.indices[np.asarray(pos)]


def _reference_epoch(indices_list, batch_size, shuffle=True, drop_last=True):
    loaders = []
    for idx in indices_list:
        ds = _Rows(idx)
        sampler = BatchSampler(RandomSampler(ds) if shuffle else SequentialSampler(ds), batch_size=batch_size, drop_last=drop_last)
        loaders.append(DataLoader(ds, sampler=sampler, batch_size=None, collate_fn=lambda b: b))
    lens = [len(dl) for dl in loaders]
    largest = loaders[int(np.argmax(lens))]
    iter_list = [cycle(dl) if dl != largest else dl for dl in loaders]  # reference _concat_dataloader.py:108-110
    return [[np.asarray(x) for x in step] for step in zip(*iter_list)]


@pytest.mark.parametrize("sizes,batch", [((1000, 1000), 128), ((900, 2100), 256), ((2100, 900), 64), ((513, 512), 512)])
def test_epoch_batches_match_torch_dataloader_stack(sizes, batch):
    from spvipes_b200.trainer import epoch_batches
    rs = np.random.RandomState(3)
    indices_list = [rs.permutation(5000)[:n] for n in sizes]
    for epoch_seed in (0, 1, 12345):
        torch.manual_seed(epoch_seed)
        want = _reference_epoch(indices_list, batch)
        want2 = _reference_epoch(indices_list, batch)  # a second epoch continues the same global RNG stream
        torch.manual_seed(epoch_seed)
        got = epoch_batches(indices_list, batch)
        got2 = epoch_batches(indices_list, batch)
        for w, g in ((want, got), (want2, got2)):
            assert len(w) == len(g) == max(n // batch for n in sizes)
            for sw, sg in zip(w, g):
                for a, b in zip(sw, sg):
                    assert np.array_equal(a, b)


def test_split_matches_reference_semantics():
    from spvipes_b200.trainer import split_indices, validate_data_split
    groups = [np.arange(0, 1003), np.arange(1003, 1003 + 777)]
    train, val, test = split_indices(groups, train_size=0.9, validation_size=None, seed=0)
    rs = np.random.RandomState(seed=0)  # reference data/_multi_datasplitter.py:66-79
    for g, gi in enumerate(groups):
        n_train, n_val = math.ceil(0.9 * len(gi)), len(gi) - math.ceil(0.9 * len(gi))
        assert validate_data_split(len(gi), 0.9, None) == (n_train, n_val)
        perm = rs.permutation(gi)
        assert np.array_equal(val[g], perm[:n_val])
        assert np.array_equal(train[g], perm[n_val:n_val + n_train])
        assert len(test[g]) == 0
