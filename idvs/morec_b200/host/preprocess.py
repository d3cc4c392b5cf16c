"""TSV -> training structures, with the reference's function names, arguments and return values
(inbatch_sasrec_e2e_text/data_utils/preprocess.py; vision: inbatch_sasrec_e2e_vision/data_utils/preprocess.py).

Semantics kept exactly: users with fewer than `min_seq_len` interactions are dropped;
sequences are cut to their last `max_seq_len + 3` items; items never seen are removed and the rest renumbered 1..N in
file order; train = all but the last two items, valid = the last `max_seq_len + 2` without the final item, test = the
last `max_seq_len + 1`; popularity p_i = (train count_i)^1.0 / sum, with the sentinel p[0] = 1.
The per-user dictionaries are the reference's return types; `pack_sequences` turns them into the flat int32 arrays
the device-side batcher (host/dataset.py) keeps in HBM.
"""
import numpy as np
import torch


def read_news(news_path):
    """ID tower: item name -> id in file order; the returned content dict maps id -> name (+ one mask sentence)"""
    id2dic, name2id, id2name = {}, {}, {}
    with open(news_path, "r") as f:
        for k, line in enumerate(f, start=1):
            doc_name = line.rstrip('\n').split('\t')[0]
            name2id[doc_name] = k
            id2dic[k] = doc_name
            id2name[k] = doc_name
    id2dic[len(id2name) + 1] = 'this is a mask sentence'
    return id2dic, name2id, id2name


def read_news_bert(news_path, args, tokenizer):
    """text tower: id -> [title, abstract, body] tokenizer outputs (max_length padding / truncation as the reference)"""
    id2dic, name2id, id2name = {}, {}, {}

    def tok(text, n):
        return tokenizer(text.lower(), max_length=n, padding='max_length', truncation=True)

    with open(news_path, "r") as f:
        for k, line in enumerate(f, start=1):
            doc_name, title, abstract = line.rstrip('\n').split('\t')[:3]
            t = tok(title, args.num_words_title) if 'title' in args.news_attributes else []
            a = tok(abstract, args.num_words_abstract) if 'abstract' in args.news_attributes else []
            b = tok(abstract[:2000], args.num_words_body) if 'body' in args.news_attributes else []   # (no body column in the TSVs)
            name2id[doc_name] = k
            id2name[k] = doc_name
            id2dic[k] = [t, a, b]
    return id2dic, name2id, id2name


def get_doc_input_bert(item_id_to_content, args):
    """-> (title ids, title mask, abstract ids, abstract mask, body ids, body mask), each int32 [N+1, n_words] or None;
    row 0 (the pad item) is all zero"""
    n = len(item_id_to_content) + 1
    out = []
    for j, (attr, width) in enumerate((('title', args.num_words_title), ('abstract', args.num_words_abstract),
                                       ('body', args.num_words_body))):
        if attr not in args.news_attributes:
            out += [None, None]
            continue
        ids = np.zeros((n, width), dtype='int32')
        msk = np.zeros((n, width), dtype='int32')
        for item_id in range(1, n):
            enc = item_id_to_content[item_id][j]
            ids[item_id] = enc['input_ids']
            msk[item_id] = enc['attention_mask']
        out += [ids, msk]
    return tuple(out)


def read_images(images_path):
    """vision: item name -> id in file order; content dict maps id -> LMDB key (the name, utf-8 encoded)"""
    id2keys, name2id, id2name = {}, {}, {}
    with open(images_path, "r") as f:
        for k, line in enumerate(f, start=1):
            name = line.rstrip('\n').split('\t')[0]
            name2id[name] = k
            id2name[k] = name
            id2keys[k] = u'{}'.format(name).encode('ascii')
    return id2keys, name2id, id2name


def _read_behaviors(behaviors_path, before_item_id_to_dic, before_item_name_to_id, before_item_id_to_name, max_seq_len,
                    min_seq_len, Log_file):
    Log_file.info("##### news number {} {} (before clearing)#####".format(len(before_item_id_to_dic), len(before_item_name_to_id)))
    Log_file.info("##### min seq len {}, max seq len {}#####".format(min_seq_len, max_seq_len))
    n_before = len(before_item_name_to_id)
    counts = np.zeros(n_before + 1, dtype=np.int64)
    user_seqs = {}
    Log_file.info('rebuild user seqs...')
    with open(behaviors_path, "r") as f:
        for line in f:
            parts = line.rstrip('\n').split('\t')
            names = parts[1].split(' ')
            if len(names) < min_seq_len:
                continue
            seq = [before_item_name_to_id[i] for i in names[-(max_seq_len + 3):]]
            user_seqs[parts[0]] = seq
            np.add.at(counts, seq, 1)
    Log_file.info("##### pairs_num {}".format(int(counts.sum())))
    kept = np.flatnonzero(counts[1:] != 0) + 1                       # before-ids that occur, in file order
    remap = np.zeros(n_before + 1, dtype=np.int64)
    remap[kept] = np.arange(1, kept.size + 1)
    item_num = int(kept.size)
    item_id_to_dic = {int(remap[b]): before_item_id_to_dic[int(b)] for b in kept}
    item_name_to_id = {before_item_id_to_name[int(b)]: int(remap[b]) for b in kept}
    Log_file.info("##### items after clearing {}, {}, {} #####".format(item_num, item_num, len(item_id_to_dic)))
    users_train, users_valid, users_test, hist_valid, hist_test = {}, {}, {}, {}, {}
    neg_sampling_list = []
    train_counts = np.zeros(item_num + 1, dtype=np.float64)
    for uid, seq in enumerate(user_seqs.values()):
        s = remap[seq].tolist()
        users_train[uid] = s[:-2]
        users_valid[uid] = s[-(max_seq_len + 2):-1]
        users_test[uid] = s[-(max_seq_len + 1):]
        np.add.at(train_counts, s[:-2], 1.0)
        neg_sampling_list.extend(s)
        hist_valid[uid] = torch.LongTensor(np.array(s[:-2]))
        hist_test[uid] = torch.LongTensor(np.array(s[:-1]))
    powered = np.power(train_counts[1:], 1.0)
    pop_prob_list = np.append([1], powered / powered.sum())
    Log_file.info("##### user seqs after clearing {}, {}, {}, {}, {}#####".format(
        len(user_seqs), len(user_seqs), len(users_train), len(users_valid), len(users_test)))
    return (item_num, item_id_to_dic, users_train, users_valid, users_test, hist_valid, hist_test, item_name_to_id,
            neg_sampling_list, pop_prob_list)


def read_behaviors(behaviors_path, before_item_id_to_dic, before_item_name_to_id, before_item_id_to_name, max_seq_len,
                   min_seq_len, Log_file):
    """text package signature (preprocess.py:5): 9 return values"""
    r = _read_behaviors(behaviors_path, before_item_id_to_dic, before_item_name_to_id, before_item_id_to_name,
                        max_seq_len, min_seq_len, Log_file)
    return r[:8] + (r[9],)


def read_behaviors_vision(behaviors_path, before_item_id_to_keys, before_item_name_to_id, before_item_id_to_name,
                          max_seq_len, min_seq_len, Log_file):
    """vision package signature: 10 return values (adds neg_sampling_list before pop_prob_list)"""
    return _read_behaviors(behaviors_path, before_item_id_to_keys, before_item_name_to_id, before_item_id_to_name,
                           max_seq_len, min_seq_len, Log_file)


def pack_sequences(u2seq):
    """{user: [item ids]} -> (flat int32 items, int32 offsets [U+1]) in user-id order"""
    n = len(u2seq)
    lens = np.fromiter((len(u2seq[u]) for u in range(n)), dtype=np.int64, count=n)
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    flat = np.empty(int(ptr[-1]), dtype=np.int32)
    for u in range(n):
        flat[ptr[u]:ptr[u + 1]] = u2seq[u]
    return flat, ptr.astype(np.int32)


def synthetic_dataset(n_users, n_items, max_seq_len, seed=12345):
    """MIND-shape synthetic interaction log: Zipf(1.0) item ids, raw lengths 5..max_seq_len+3 with 35 %
    at the cap; returns what read_behaviors returns (item contents are up to the caller)"""
    g = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, n_items + 1, dtype=np.float64)
    w /= w.sum()
    perm = g.permutation(n_items) + 1
    cap = max_seq_len + 3
    lens = np.where(g.random(n_users) < 0.35, cap, g.integers(5, cap + 1, size=n_users))
    users_train, users_valid, users_test, hist_valid, hist_test = {}, {}, {}, {}, {}
    train_counts = np.zeros(n_items + 1, dtype=np.float64)
    neg = []
    draws = perm[g.choice(n_items, size=int(lens.sum()), p=w)]
    off = 0
    for u in range(n_users):
        s = draws[off:off + lens[u]].tolist()
        off += lens[u]
        users_train[u], users_valid[u], users_test[u] = s[:-2], s[-(max_seq_len + 2):-1], s[-(max_seq_len + 1):]
        np.add.at(train_counts, s[:-2], 1.0)
        neg.extend(s)
        hist_valid[u] = torch.LongTensor(np.array(s[:-2]))
        hist_test[u] = torch.LongTensor(np.array(s[:-1]))
    train_counts[1:] += 1e-3                                   # keep p > 0 for items that only occur as targets
    pop = np.append([1], train_counts[1:] / train_counts[1:].sum())
    return n_items, None, users_train, users_valid, users_test, hist_valid, hist_test, None, neg, pop
