"""SURVEY.md 8(a) row a3 on the device: ``KGTensors.from_dir`` over a reference-style data directory (triples, numeric
literal files, text-literal pickles) -> device tables, initial Laplacian ``A_in``, and the gate / embedding pass fed by
them, against the oracle on the same files."""
import argparse
import pickle

import numpy as np
import pytest
import torch

import literalkg_oracle as O

pytestmark = pytest.mark.gpu


def test_from_dir_tables_on_device_feed_the_path(tmp_path):
    import literalkg_b200 as L
    from literalkg_b200 import dataloader as D
    rng = np.random.default_rng(11)
    n, dim = 300, 8
    trip = np.unique(np.stack([rng.integers(0, n, 3000), rng.integers(0, 5, 3000), rng.integers(0, n, 3000)], 1), axis=0)
    rng.shuffle(trip)
    (tmp_path / "pre_training_train.txt").write_text("\n".join(f"{h} {r} {t}" for h, r, t in trip) + "\n")
    age = "".join(f"{i}\t{float(rng.integers(1, 20))}\n" for i in rng.choice(n, 40, replace=False))
    wgt = "".join(f"{i}\t{float(rng.integers(1, 90))}\n" for i in rng.choice(n + 20, 60, replace=False))   # ids past the KG
    (tmp_path / "age_dict.txt").write_text("40\n" + age)
    (tmp_path / "weight_dict.txt").write_text("60\n" + wgt)
    for name in ("cc_dict.pickle", "treatment_dict.pickle"):
        with open(tmp_path / name, "wb") as fh:
            pickle.dump({int(i): rng.normal(size=dim).astype(np.float32) * 0.3 for i in rng.choice(n + 20, 50, replace=False)}, fh)
    kt = L.KGTensors.from_dir(str(tmp_path), numeric_dim=2, text_dim=dim, device="cuda")
    _, num, txt = D.read_data_dir(str(tmp_path), numeric_dim=2, text_dim=dim)
    assert kt.num_embedding_table.is_cuda and kt.text_embedding_table.is_cuda
    assert kt.n_entities == num.shape[0] == txt.shape[0] and kt.n_entities > n          # literal ids extend the id space
    assert np.array_equal(kt.num_embedding_table.cpu().numpy(), num) and np.array_equal(kt.text_embedding_table.cpu().numpy(), txt)
    want = O.numeric_literal_table([(tmp_path / f).read_text() for f in ("age_dict.txt", "weight_dict.txt")], kt.n_entities, 2)
    rows_txt = np.array(sorted(set().union(*[pickle.load(open(tmp_path / f, "rb")).keys()
                                             for f in ("cc_dict.pickle", "treatment_dict.pickle")])))
    want[rows_txt] = 0                                                               # dataloader.py:147-150
    assert np.array_equal(num, want)
    li, lv = O.laplacian_A_in(trip[:, 0], trip[:, 2], trip[:, 1], kt.n_entities)
    assert np.array_equal(kt.A_in.indices().cpu().numpy(), li) and np.array_equal(kt.A_in.values().cpu().numpy(), lv)
    cfg = O.OracleConfig(embed_dim=12, relation_dim=12, scale_gat_dim=16, txt_lit_dim=dim, conv_dim=8, n_conv_layers=2,
                         mess_dropout=0.0)
    p = O.init_params(cfg, kt.n_entities, kt.n_relations, seed=5)
    p["entity_embed.weight"] *= 10
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m = L.LiteralKG(args, kt.n_entities, kt.n_relations, kt.A_in, kt.num_embedding_table, kt.text_embedding_table)
    m.load_state_dict(p, strict=False)
    m = m.cuda().eval()
    with torch.no_grad():
        h0 = m.gate_embeddings()
        out = m.gat_embeddings()
    ref_h0 = O.gate_embeddings(p, cfg, torch.from_numpy(num), torch.from_numpy(txt))
    ref = O.gat_embeddings(p, cfg, torch.from_numpy(li), torch.from_numpy(lv), torch.from_numpy(num), torch.from_numpy(txt))
    rel = lambda a, b: ((a.cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()
    assert rel(h0, ref_h0) < 1e-5 and rel(out, ref) < 1e-3
