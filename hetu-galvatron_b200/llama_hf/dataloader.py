"""Synthetic token data for the Llama family -- same generator, seeds and collate contract as
``galvatron/models/llama_hf/dataloader.py:40-80`` (``DataLoaderForLlama`` + ``random_collate_fn``): per-sample length
``randint(1, seq+1)``, tokens uniform in [0, vocab) zero-padded to seq+1, tokens = x[:, :-1], labels = x[:, 1:]."""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset


def set_seed(seed=1234):
    """galvatron/utils/training_utils.py:7-11"""
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def random_collate_fn(batch):
    tokens_ = torch.stack(batch, dim=0)
    labels = tokens_[:, 1:].contiguous()
    tokens = tokens_[:, :-1].contiguous()
    return tokens, {"attention_mask": None, "labels": labels}, None


class DataLoaderForLlama(Dataset):
    def __init__(self, args, device, dataset_size=2560 * 16):
        self.vocab_size, self.sentence_length, self.dataset_size, self.device = args.vocab_size, args.seq_length, dataset_size, device
        self.data_length = np.random.randint(1, self.sentence_length + 1, (dataset_size,))
        ids = np.random.randint(0, self.vocab_size, (dataset_size, self.sentence_length + 1))
        ids[np.arange(self.sentence_length + 1)[None, :] >= self.data_length[:, None]] = 0
        self.input_ids = ids

    def __len__(self):
        return self.dataset_size

    def __getitem__(self, idx):
        if idx >= self.dataset_size:
            raise IndexError
        return torch.from_numpy(self.input_ids[idx]).long().to(self.device)


def get_train_loader(args, device, dataset_size, rank_in_dp, dp_size):
    """Each data-parallel rank reads a disjoint slice (galvatron/utils/training_utils.py:13 distributed_dataloader)."""
    ds = DataLoaderForLlama(args, device, dataset_size)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=dp_size, rank=rank_in_dp, shuffle=False)
    return DataLoader(ds, batch_size=args.global_train_batch_size // dp_size, sampler=sampler, collate_fn=random_collate_fn)
