"""Drop-in packages `inbatch_sasrec_e2e_text/` and `inbatch_sasrec_e2e_vision/`, CPU side: the
reference launchers' command lines parse, the data layer reproduces the reference's structures, and the device-side
batcher produces exactly the batches the reference's Dataset + default collate would."""
import importlib
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

# the flag list train_bert_base.py:40-50 and train_swin_tiny.py:40-49 pass to run.py (values as those launchers format them)
TEXT_CMD = ("run.py --root_data_dir ../../ --dataset Dataset/MIND-large --behaviors mind_60w_users.tsv --news mind_60w_items.tsv "
            "--mode train --item_tower modal --load_ckpt_name None --label_screen modal_bs128_ed512_lr0.0001_dp0.1_L20.01_Flr5e-05 "
            "--logging_num 4 --testing_num 1 --l2_weight 0.01 --fine_tune_l2_weight 0.01 --drop_rate 0.1 --batch_size 128 "
            "--lr 0.0001 --embedding_dim 512 --news_attributes title --bert_model_load bert_base_uncased --epoch 300 "
            "--freeze_paras_before 0 --fine_tune_lr 5e-05")
VISION_CMD = ("run.py --root_data_dir ../../ --dataset Dataset/HM-large --behaviors hm_50w_users.tsv --images hm_50w_items.tsv "
              "--lmdb_data hm_50w_items.lmdb --mode train --item_tower modal --load_ckpt_name None --label_screen x "
              "--logging_num 4 --testing_num 1 --l2_weight 0.1 --fine_tune_l2_weight 0.1 --drop_rate 0.1 --batch_size 64 "
              "--lr 0.0001 --embedding_dim 2048 --CV_resize 224 --CV_model_load swin_tiny --epoch 200 "
              "--freeze_paras_before 0 --fine_tune_lr 0.0001")


def _pkg_parse(pkg, argv):
    code = ("import json, sys; import parameters; a = parameters.parse_args(sys.argv[1:]); "
            "print(json.dumps({k: v for k, v in vars(a).items()}))")
    r = subprocess.run([sys.executable, "-c", code] + argv, cwd=os.path.join(ROOT, pkg), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    import json
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("spelling", ["--local_rank=1", "--local-rank=1"])
def test_text_launcher_command_line_parses(spelling):
    a = _pkg_parse("inbatch_sasrec_e2e_text", [spelling] + TEXT_CMD.split()[1:])
    assert a["news"] == "mind_60w_items.tsv" and a["news_attributes"] == ["title"]      # --news is a real flag now
    assert a["local_rank"] == 1 and a["batch_size"] == 128 and a["bert_model_load"] == "bert_base_uncased"
    assert a["fine_tune_lr"] == 5e-05 and a["freeze_paras_before"] == 0 and a["embedding_dim"] == 512


def test_vision_launcher_command_line_parses_and_local_rank_env(monkeypatch):
    monkeypatch.setenv("LOCAL_RANK", "3")
    a = _pkg_parse("inbatch_sasrec_e2e_vision", VISION_CMD.split()[1:])
    assert a["local_rank"] == 3 and a["CV_model_load"] == "swin_tiny" and a["lmdb_data"] == "hm_50w_items.lmdb"
    assert a["embedding_dim"] == 2048 and a["max_seq_len"] == 10


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference launchers (authoring container)")
@pytest.mark.parametrize("pkg,launcher", [("inbatch_sasrec_e2e_text", "train_bert_base.py"), ("inbatch_sasrec_e2e_text", "train_id.py"),
                                          ("inbatch_sasrec_e2e_vision", "train_swin_tiny.py")])
def test_unmodified_reference_launcher_drives_our_parser(pkg, launcher, monkeypatch):
    """execute the reference's launcher script UNEDITED with os.system captured; every command it would start must be
    accepted by this repo's parameters.py (after the `python -m torch.distributed.launch ... run.py` prefix)"""
    cmds = []
    monkeypatch.setattr(os, "system", lambda c: cmds.append(c) or 0)
    src = open(os.path.join(REF, pkg, launcher)).read()
    exec(compile(src, launcher, "exec"), {"__name__": "__main__"})
    assert cmds
    for c in cmds[:2]:
        toks = c.split()
        argv = toks[toks.index("run.py") + 1:]
        a = _pkg_parse(pkg, ["--local-rank=0"] + argv)
        assert a["mode"] == "train" and a["local_rank"] == 0


def test_packages_export_reference_surface():
    for pkg, names in [("inbatch_sasrec_e2e_text", ["read_news", "read_news_bert", "get_doc_input_bert", "read_behaviors",
                                                    "BuildTrainDataset", "BuildEvalDataset", "SequentialDistributedSampler",
                                                    "eval_model", "get_item_embeddings", "setuplogger", "save_model",
                                                    "para_and_log", "get_checkpoint", "report_time_train"]),
                       ("inbatch_sasrec_e2e_vision", ["read_images", "read_behaviors", "Build_Lmdb_Dataset", "Build_Id_Dataset",
                                                      "LMDB_Image", "eval_model", "get_itemId_embeddings",
                                                      "get_itemLMDB_embeddings", "SequentialDistributedSampler"])]:
        code = ("import data_utils, model, run, parameters; import data_utils.utils as u; "
                f"missing = [n for n in {names!r} if not hasattr(data_utils, n)]; assert not missing, missing; "
                "assert hasattr(model, 'Model') and hasattr(run, 'train') and hasattr(run, 'run_eval') and hasattr(run, 'setup_seed'); "
                "assert all(hasattr(u, n) for n in ('os', 'time', 'torch', 'argparse', 'setuplogger'))")
        r = subprocess.run([sys.executable, "-c", code], cwd=os.path.join(ROOT, pkg), capture_output=True, text=True)
        assert r.returncode == 0, (pkg, r.stderr[-1500:])     # (vision imports although `lmdb` is not installed)


def _write_tsv(tmp_path, n_users=400, n_items=300, seed=0):
    g = np.random.default_rng(seed)
    words = ["alpha", "beta", "gamma", "delta", "news", "today", "market", "sports", "the", "of", "rain", "city"]
    items = tmp_path / "items.tsv"
    with open(items, "w") as f:
        for i in range(n_items):
            t = " ".join(g.choice(words, size=g.integers(3, 9)))
            f.write(f"N{i}\t{t}\t{t} abstract\n")
    users = tmp_path / "users.tsv"
    with open(users, "w") as f:
        for u in range(n_users):
            n = int(g.integers(2, 30))
            f.write(f"U{u}\t" + " ".join(f"N{int(x)}" for x in g.integers(0, n_items, size=n)) + "\n")
    return str(items), str(users)


def _ref_module(name):
    spec = importlib.util.spec_from_file_location("_ref_pp_" + name, os.path.join(REF, "inbatch_sasrec_e2e_text", "data_utils", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Log:
    def info(self, *a, **k):
        pass


@pytest.mark.skipif(not os.path.isdir(REF), reason="compares against the reference's own preprocess.py (authoring container)")
@pytest.mark.parametrize("source", ["synthetic_tsv", "real_mind_head"])
def test_preprocess_matches_reference(tmp_path, source):
    """read_news / read_behaviors: identical item numbering, splits, histories and popularity to the reference's
    implementation (data_utils/preprocess.py:5-98) on synthetic TSVs and on the head of the bundled MIND files"""
    from idvs.morec_b200.host import preprocess as PP
    R = _ref_module("preprocess")
    if source == "synthetic_tsv":
        items, users = _write_tsv(tmp_path)
    else:
        items = os.path.join(REF, "dataset", "MIND", "mind_60w_items.tsv")
        users = str(tmp_path / "users_head.tsv")
        with open(os.path.join(REF, "dataset", "MIND", "mind_60w_users.tsv")) as f, open(users, "w") as o:
            for i, line in enumerate(f):
                if i >= 3000:
                    break
                o.write(line)
    for L, mn in [(20, 5), (10, 3)]:
        a = R.read_behaviors(users, *R.read_news(items), L, mn, _Log())
        b = PP.read_behaviors(users, *PP.read_news(items), L, mn, _Log())
        assert a[0] == b[0] and a[1] == b[1] and a[7] == b[7]                # item_num, id -> content, name -> id
        for k in (2, 3, 4):                                                 # users_train / valid / test
            assert a[k] == b[k]
        for k in (5, 6):                                                    # histories (LongTensors)
            assert a[k].keys() == b[k].keys() and all(torch.equal(a[k][u], b[k][u]) for u in a[k])
        assert np.array_equal(np.asarray(a[8], dtype=np.float64), np.asarray(b[8], dtype=np.float64))


def test_device_batcher_equals_reference_dataset_and_sampler():
    """DeviceBatcher.batch == default_collate of BuildTrainDataset samples (dataset.py:24-36), and its epoch order ==
    torch's DistributedSampler (run.py:115, 229)"""
    from idvs.morec_b200.host.dataset import BuildTrainDataset, DeviceBatcher, SequentialDistributedSampler
    from torch.utils.data import default_collate
    from torch.utils.data.distributed import DistributedSampler
    g = np.random.default_rng(5)
    L, N, T = 7, 50, 6
    u2seq = {u: [int(x) for x in g.integers(1, N + 1, size=g.integers(3, L + 2))] for u in range(37)}
    content = g.integers(0, 1000, size=(N + 1, 2 * T)).astype(np.int32)
    content[0] = 0
    for modal in (True, False):
        ds = BuildTrainDataset(u2seq, content if modal else np.arange(N + 1), N, L, modal)
        db = DeviceBatcher(u2seq, content if modal else None, L, modal, "cpu")
        users = torch.tensor([3, 0, 36, 17, 17])
        ref = default_collate([ds[int(u)] for u in users])
        got = db.batch(users)
        for r, o in zip(ref, got):
            assert r.dtype == o.dtype and torch.equal(r, o)
        assert np.array_equal(db.host_ids(users.numpy()), ref[0].numpy())
    for world in (1, 2, 4):
        for rank in range(world):
            s = DistributedSampler(range(37), num_replicas=world, rank=rank)
            s.set_epoch(3)
            assert list(s) == db.epoch_order(3, rank, world).tolist()
    s = SequentialDistributedSampler(range(37), batch_size=8, rank=1, num_replicas=2)
    assert len(s) == 24 and list(s)[:3] == [24, 25, 26] and list(s)[-1] == 36


def test_synthetic_dataset_shapes():
    from idvs.morec_b200.host.preprocess import pack_sequences, synthetic_dataset
    item_num, _, tr, va, te, hv, ht, _, neg, pop = synthetic_dataset(200, 150, 10)
    assert item_num == 150 and len(tr) == len(va) == len(te) == 200 and abs(pop[1:].sum() - 1) < 1e-9 and pop[0] == 1
    assert all(3 <= len(tr[u]) <= 11 and len(va[u]) <= 11 and va[u][:-1] == tr[u][-(len(va[u]) - 1):] for u in tr)
    flat, ptr = pack_sequences(tr)
    assert ptr[-1] == flat.size == sum(len(v) for v in tr.values())
