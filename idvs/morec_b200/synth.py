"""Seeded synthetic batches of SURVEY.md §8(d) (integer-exact; shared by bench.py, the tests and the oracle).

Mirrors the input layout the reference's data layer produces (inbatch_sasrec_e2e_text/data_utils/dataset.py:24-36:
left-padded item sequences + log_mask) and its popularity table (data_utils/preprocess.py:71-76).
"""
import numpy as np
import torch


def log_mask_from_ids(ids: torch.Tensor) -> torch.Tensor:
    """dataset.py:24-36: log_mask[b,t] = 1 <=> slot t (t < L) of user b holds a real item."""
    return (ids[:, :-1] != 0).to(torch.float32)


def synth_batch(B: int, L: int, N: int, T: int, seed: int, *, modal: bool, mind_shape: bool = True,
                vocab_lo: int = 1000, vocab_hi: int = 30000, n_users_pop: int = 20000):
    """MIND-shape synthetic batch: Zipf(1.0) item ids over 1..N, 35 % full-length users and the rest
    uniform on [3, L] items, left padded; titles of uniform length [6, T] with [CLS]=101 first;
    popularity p_i computed from a synthetic train split exactly as preprocess.py:71-76 (count / total,
    p[0] = 1), with every in-batch id guaranteed p > 0.
    Returns dict(ids [B,L+1] i64, items [C,2T] | [C] i64, log_mask [B,L] f32, pop_prob [N+1] f64,
    item_content [N+1, 2T] i64 | None).
    """
    g = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, N + 1, dtype=np.float64)
    w /= w.sum()
    perm = g.permutation(N) + 1                       # popularity rank -> item id

    def draw(n):
        return perm[g.choice(N, size=n, p=w)]

    counts = np.zeros(N + 1, dtype=np.float64)
    pop_draw = draw(n_users_pop * 8)
    np.add.at(counts, pop_draw, 1.0)
    ids = np.zeros((B, L + 1), dtype=np.int64)
    for b in range(B):
        if (not mind_shape) or g.random() < 0.35:
            n = L + 1
        else:
            n = int(g.integers(3, L + 1))
        seq = draw(n)
        ids[b, L + 1 - n:] = seq
        np.add.at(counts, seq, 1.0)
    pop = counts[1:] / counts[1:].sum()
    pop_prob = np.append([1.0], pop)
    out = dict(ids=torch.from_numpy(ids), log_mask=log_mask_from_ids(torch.from_numpy(ids)),
               pop_prob=torch.from_numpy(pop_prob), counts=torch.from_numpy(counts))
    if modal:
        content = np.zeros((N + 1, 2 * T), dtype=np.int64)
        lens = g.integers(min(6, T), T + 1, size=N + 1)
        tok = g.integers(vocab_lo, vocab_hi, size=(N + 1, T))
        ar = np.arange(T)[None, :]
        am = (ar < lens[:, None]).astype(np.int64)
        tok = tok * am
        tok[:, 0] = 101
        content[:, :T] = tok
        content[:, T:] = am
        content[0] = 0
        out["item_content"] = torch.from_numpy(content)
        out["items"] = torch.from_numpy(content[ids.reshape(-1)])
    else:
        out["item_content"] = None
        out["items"] = torch.from_numpy(ids.reshape(-1).copy())
    return out


def pop_from_batches(batches):
    """ONE popularity table for a set of synthetic batches (a model is built with a single pop_prob_list): the summed
    train-split counts of all of them, so every id of every batch has p > 0 (log p finite), p[0] = 1."""
    counts = sum(b["counts"] for b in batches)
    pop = counts[1:] / counts[1:].sum()
    return torch.cat([torch.ones(1, dtype=pop.dtype), pop])


def synth_catalogue(N: int, T: int, seed: int, vocab_lo: int = 1000, vocab_hi: int = 30000):
    """ONE synthetic catalogue `item_content` int64 [N+1, 2T] (T token ids || T attention mask, row 0 = pad item): titles
    of uniform length [6, T] with [CLS]=101 first -- the array run.py:93-98 builds from the news file.  Batches that
    belong to one run must index the SAME catalogue (an item has one title, on every rank)."""
    g = np.random.default_rng(seed)
    content = np.zeros((N + 1, 2 * T), dtype=np.int64)
    lens = g.integers(min(6, T), T + 1, size=N + 1)
    tok = g.integers(vocab_lo, vocab_hi, size=(N + 1, T))
    am = (np.arange(T)[None, :] < lens[:, None]).astype(np.int64)
    tok = tok * am
    tok[:, 0] = 101
    content[:, :T] = tok
    content[:, T:] = am
    content[0] = 0
    return torch.from_numpy(content)
