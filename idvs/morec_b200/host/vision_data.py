"""Image side of the vision package's data layer (inbatch_sasrec_e2e_vision/data_utils/dataset.py:16-99): `LMDB_Image`,
`Build_Lmdb_Dataset`, `Build_Id_Dataset` with the reference's names, constructor arguments and sample layout.

`lmdb` is imported LAZILY, inside the classes that open a database: the package is not installed in every
environment (it is absent from this image) and the reference imports it at module level
(inbatch_sasrec_e2e_vision/data_utils/dataset.py:10), which makes even the ID-tower run fail to import.
"""
import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset

from . import preprocess as PP


def _lmdb():
    try:
        import lmdb
    except ImportError as e:                       # pragma: no cover - depends on the environment
        raise RuntimeError("the image store is an LMDB database: `pip install lmdb` (only needed for --item_tower modal "
                           "with real images; use --synthetic users,items for a synthetic run)") from e
    return lmdb


class LMDB_Image:
    """pickled record of the image store: raw uint8 pixels + shape + id"""

    def __init__(self, image, id):
        self.channels = image.shape[2]
        self.size = image.shape[:2]
        self.image = image.tobytes()
        self.id = id

    def get_image(self):
        return np.frombuffer(self.image, dtype=np.uint8).reshape(*self.size, self.channels)


def _left_pad(seq, width):
    pad = width - len(seq)
    ids = torch.zeros(width, dtype=torch.long)
    ids[pad:] = torch.as_tensor(seq, dtype=torch.long)
    log_mask = torch.zeros(width - 1, dtype=torch.float32)
    log_mask[pad:] = 1.0
    return ids, log_mask, pad


class Build_Id_Dataset(Dataset):
    def __init__(self, u2seq, item_num, max_seq_len, neg_sampling_list):
        self.u2seq, self.item_num, self.max_seq_len = u2seq, item_num, max_seq_len + 1
        self.neg_sampling_list = neg_sampling_list

    def __len__(self):
        return len(self.u2seq)

    def __getitem__(self, user_id):
        ids, log_mask, _ = _left_pad(self.u2seq[user_id], self.max_seq_len)
        return ids, ids, log_mask


class ImageStore:
    """read-only LMDB of LMDB_Image records -> normalised float tensors [3, R, R] (Resize, ToTensor, Normalize(0.5, 0.5))"""

    def __init__(self, db_path, resize):
        lmdb = _lmdb()
        self.env = lmdb.open(db_path, subdir=os.path.isdir(db_path), readonly=True, lock=False, readahead=False,
                             meminit=False)
        self.resize = resize

    def get(self, txn, key):
        from PIL import Image
        rec = pickle.loads(txn.get(key))
        img = Image.fromarray(rec.get_image()).convert('RGB').resize((self.resize, self.resize), Image.BILINEAR)
        x = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255.0)
        return x.sub_(0.5).div_(0.5)


class Build_Lmdb_Dataset(Dataset):
    def __init__(self, u2seq, item_num, max_seq_len, db_path, item_id_to_keys, resize, neg_sampling_list):
        self.u2seq, self.item_num, self.max_seq_len = u2seq, item_num, max_seq_len + 1
        self.item_id_to_keys, self.resize = item_id_to_keys, resize
        self.neg_sampling_list = neg_sampling_list
        self.store = ImageStore(db_path, resize)

    def __len__(self):
        return len(self.u2seq)

    def __getitem__(self, user_id):
        seq = self.u2seq[user_id]
        ids, log_mask, pad = _left_pad(seq, self.max_seq_len)
        items = torch.zeros(self.max_seq_len, 3, self.resize, self.resize)       # pad slots: the all-zero image
        with self.store.env.begin() as txn:
            for i, it in enumerate(seq):
                items[pad + i] = self.store.get(txn, self.item_id_to_keys[it])
        return ids, items, log_mask


def load_vision_data(args, use_modal, Log_file):
    """-> (item_num, item_content, users_train, users_valid, users_history_for_valid, pop_prob_list); item_content is a
    float16 image tensor [N+1, 3, R, R] for synthetic runs (kept on the device by the trainer), else the LMDB key map"""
    if 'None' not in args.synthetic:
        n_users, n_items = (int(x) for x in args.synthetic.split(','))
        (item_num, _, users_train, users_valid, users_test, hist_valid, hist_test, _, _, pop) = \
            PP.synthetic_dataset(n_users, n_items, args.max_seq_len)
        content = None
        if use_modal:
            g = torch.Generator().manual_seed(4243)
            content = torch.randn(n_items + 1, 3, args.CV_resize, args.CV_resize, generator=g).to(torch.float16)
            content[0] = 0
        return item_num, content, users_train, users_valid, hist_valid, pop
    b_keys, b_n2i, b_i2n = PP.read_images(os.path.join(args.root_data_dir, args.dataset, args.images))
    (item_num, item_id_to_keys, users_train, users_valid, users_test, hist_valid, hist_test, item_name_to_id, neg, pop) = \
        PP.read_behaviors_vision(os.path.join(args.root_data_dir, args.dataset, args.behaviors), b_keys, b_n2i, b_i2n,
                                 args.max_seq_len, args.min_seq_len, Log_file)
    return item_num, item_id_to_keys, users_train, users_valid, hist_valid, pop
