"""CPU: host-side logic of the drivers -- text front-end, batching, utterance sharding (incl. a
world_size-2 gloo run of the N>1 path)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import ttsmodel_oracle as O
from oracle import weights as W
from spoofsv_b200 import text as T
from spoofsv_b200.synth import Unit, corpus_units, plan_batches, shard_range


def test_text2id_agrees_with_oracle_on_all_lines():
    _, _, lines = W.load_fixtures()
    enc = T.encode_lines(lines)
    for s, ids in zip(lines, enc):
        assert np.array_equal(ids, O.text2id(s))
    assert T.vocab_len() == 34


def test_pad_batch():
    rows = [np.array([[3, 4, 1]]), np.array([[5, 1]])]
    out = T.pad_batch(rows)
    assert out.shape == (2, 3) and out.dtype == np.int64 and out[1].tolist() == [5, 1, 0]
    assert T.pad_batch(rows, 58).shape == (2, 58)
    with pytest.raises(ValueError):
        T.pad_batch(rows, 2)


def test_shard_range_partitions_exactly():
    for n, w in [(77760, 8), (77760, 1), (10, 4), (3, 8), (0, 2)]:
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in (shard_range(77760, 8, r) for r in range(8))] == [9720] * 8
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_corpus_units_and_batches():
    units = corpus_units(108, 720)
    assert len(units) == 77760 and units[0] == Unit(0, 0) and units[720] == Unit(1, 0)
    batches = plan_batches(units[:130], 64)
    assert [len(b) for b in batches] == [64, 64, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_units, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_units, world, rank)
    # what bench.py does across ranks: barrier, per-rank work, MAX-reduce of the time, SUM of units
    dist.barrier()
    mine = torch.zeros(n_units, dtype=torch.int64)
    mine[lo:hi] = 1
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save({"cover": mine, "tmax": t, "count": cnt}, os.path.join(out_dir, "r0.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    import torch.multiprocessing as mp
    n_units = 1001
    mp.spawn(_gloo_worker, args=(2, _free_port(), n_units, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(tmp_path / "r0.pt")
    assert bool((r["cover"] == 1).all())          # every unit owned by exactly one rank, no collective on data
    assert float(r["tmax"]) == 2.0 and int(r["count"]) == n_units


# --------------------------------------------------------------------------- corpus driver (CLI surface)
def test_driver_cli_flags_match_the_reference():
    from spoofsv_b200 import generate_test_utterances as G
    ps = G.build_parser()
    a = ps.parse_args(["-C", "config.json", "-T", "19-08-17_13-05-42"])            # reference defaults (:46-50)
    assert (a.train_spk_num, a.enroll_utt_num, a.eval_utt_num) == (88, 3, 20)
    a = ps.parse_args(["--configuration", "c.json", "--current_time", "x", "--eval_utt_num", "5"])
    assert a.configuration == "c.json" and a.current_time == "x" and a.eval_utt_num == 5
    with pytest.raises(SystemExit):
        ps.parse_args(["-C", "c.json"])                                             # -T is required, as in the reference


def test_driver_texts_speakers_plan(tmp_path, write_driver_cfg):
    from spoofsv_b200 import generate_test_utterances as G
    path, cfg = write_driver_cfg(tmp_path)
    cfg2 = G.load_config(str(path))
    ids = G.read_texts(cfg2, 4)
    _, _, lines = W.load_fixtures()
    want = O.pad_text_ids([O.text2id(s) for s in lines[:4]])                       # reference padding: batch maximum
    assert np.array_equal(ids, want[:, 0, :].numpy())
    with pytest.raises(ValueError):
        G.read_texts(cfg2, 7)
    assert G.list_speakers(cfg2) == ["p225", "p226", "p227"]
    (tmp_path / "no_corpus" / "wav22" / "p300").mkdir(parents=True)                 # with the corpus: its directory listing
    assert G.list_speakers(cfg2) == ["p300"]
    assert G.output_name("p225", 0) == "s225/s225_001"                              # reference wav naming (:139)
    # two ranks cover every (speaker, sentence) unit exactly once, speaker-major
    seen = [u for r in range(2) for b in G.plan(3, 4, 2, r, 5) for u in b]
    assert seen == [Unit(s, k) for s in range(3) for k in range(4)]
    assert [len(b) for b in G.plan(3, 4, 1, 0, 5)] == [5, 5, 2]


def test_collate_functions_pad_like_the_reference():
    """data/dataset.py:184-263: text ids are padded with id 0 ('P'), spectrograms with zero frames, to the batch maxima."""
    import torch
    from spoofsv_b200.data import collate_pad_2, collate_pad_3, collate_pad_4
    mk = lambda t_mel, n_txt: {"data_0": torch.ones(80, t_mel), "data_1": torch.full((1, n_txt), 5, dtype=torch.int64),
                              "data_2": torch.ones(200, 1), "data_3": torch.ones(513, 4 * t_mel)}
    b3 = collate_pad_3([mk(7, 4), mk(5, 9)])
    assert tuple(b3["data_0"].shape) == (2, 80, 7) and tuple(b3["data_1"].shape) == (2, 1, 9) and tuple(b3["data_2"].shape) == (2, 200, 1)
    assert b3["data_1"].dtype == torch.int64 and int(b3["data_1"][0, 0, 4:].sum()) == 0 and float(b3["data_0"][1, :, 5:].abs().sum()) == 0
    b4 = collate_pad_4([mk(3, 2), mk(6, 2)])
    assert tuple(b4["data_3"].shape) == (2, 513, 24) and float(b4["data_3"][0, :, 12:].abs().sum()) == 0
    b2 = collate_pad_2([{"data_0": torch.ones(80, 3), "data_1": torch.ones(513, 12)}, {"data_0": torch.ones(80, 4), "data_1": torch.ones(513, 16)}])
    assert tuple(b2["data_0"].shape) == (2, 80, 4) and tuple(b2["data_1"].shape) == (2, 513, 16)


def test_dataset_reads_the_reference_path_lists(tmp_path):
    from spoofsv_b200.data import dataset
    from spoofsv_b200 import text as T
    root = tmp_path / "d"
    (root / "data_path" / "ordinary").mkdir(parents=True)
    (root / "data_path" / "ubm-finetune").mkdir(parents=True)
    (root / "data_path" / "ordinary" / "wav.path.train").write_text("/x/wav22/p225/p225_001.wav\n/x/wav22/p226/p226_004.wav\n")
    (root / "data_path" / "ordinary" / "txt.path.train").write_text("/x/txt/p225/p225_001.txt\n/x/txt/p226/p226_004.txt\n")
    (root / "data_path" / "ubm-finetune" / "wav.path.ubm.validate").write_text("/x/wav22/p227/p227_002.wav\n")
    (root / "data_path" / "ubm-finetune" / "txt.path.ubm.validate").write_text("/x/txt/p227/p227_002.txt\n")
    cfg = {"DATA_ROOT_DIR": str(root) + "/", "SPK_EMB_DIR": "/x/spk/", "VOCABULARY": T.DEFAULT_VOCABULARY}
    ds = dataset(cfg, mode="train")
    assert len(ds) == 2 and ds.wavlist[1].endswith("p226_004.wav") and ds.wavlist[1][-12:-8] == "p226"
    assert len(dataset(cfg, mode="validate", pattern="ubm-finetune", stage="ubm")) == 1
    import pytest
    with pytest.raises(ValueError):
        dataset(cfg, mode="train", pattern="ubm-finetune", stage=None)
