"""Host side of the data layer: file readers against the oracle's restatement of the
reference loader, on synthetic files and -- where the reference checkout is mounted -- on its bundled data/Test."""
import os

import numpy as np
import pytest

import literalkg_oracle as O
from literalkg_b200 import dataloader as D

REF_DATA = os.path.join(os.environ.get("LKG_REFERENCE", "/root/reference"), "data", "Test")


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_read_triples_dedup_keeps_first_occurrence_and_order(tmp_path):
    rng = np.random.default_rng(0)
    trip = rng.integers(0, 40, size=(500, 3))
    trip[:, 1] %= 5
    trip = np.concatenate([trip, trip[rng.integers(0, 500, 80)]])       # exact duplicate rows
    rng.shuffle(trip)
    text = "\n".join(f"{h} {r} {t}" for h, r, t in trip) + "\n"
    got = D.read_triples(_write(tmp_path, "kg.txt", text))
    h, t, r = O.parse_triples(text)
    assert np.array_equal(got[:, 0], h) and np.array_equal(got[:, 1], r) and np.array_equal(got[:, 2], t)
    assert D.relation_order(got[:, 1]) == O.relation_order(r)
    # ids too wide for one packed key take the row-wise path: same result
    big = trip.astype(np.int64)
    big[:, 0] += 1 << 40
    big[:, 2] += 1 << 41
    textb = "\n".join(f"{h} {r} {t}" for h, r, t in big) + "\n"
    gotb = D.read_triples(_write(tmp_path, "kg_big.txt", textb))
    hb, tb, rb = O.parse_triples(textb)
    assert np.array_equal(gotb[:, 0], hb) and np.array_equal(gotb[:, 2], tb) and np.array_equal(gotb[:, 1], rb)


def test_numeric_literals_match_the_oracle(tmp_path):
    age = "3\n0\t4.0\n2\t9.0\n7\t1.5\n2\t3.0\n"           # leading count line, a repeated id keeps its last value
    weight = "2\n2\t10.0\n5\t40.0\n"
    paths = [_write(tmp_path, "age_dict.txt", age), _write(tmp_path, "weight_dict.txt", weight)]
    table, max_id = D.read_numeric_literals(paths, 9, 2)
    assert max_id == 7 and table.shape == (9, 2)
    assert np.array_equal(table, O.numeric_literal_table([age, weight], 9, 2))
    assert table[2, 0] == 0 and table[2, 1] == np.float32(11.0 / 40.0)     # the later file resets the whole row


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "pre_training_train.txt")),
                    reason="reference checkout not mounted")
def test_bundled_reference_data_dir():
    path = os.path.join(REF_DATA, "pre_training_train.txt")
    got = D.read_triples(path)
    h, t, r = O.parse_triples(open(path).read())
    assert np.array_equal(got[:, 0], h) and np.array_equal(got[:, 1], r) and np.array_equal(got[:, 2], t)
    import pandas as pd                                      # the reference's own load_graph (dataloader.py:186-190)
    ref = pd.read_csv(path, sep=" ", names=["h", "r", "t"], engine="python").drop_duplicates()
    assert np.array_equal(got, ref[["h", "r", "t"]].to_numpy())
    files = [os.path.join(REF_DATA, f) for f in ("age_dict.txt", "weight_dict.txt")]
    n = int(max(got[:, 0].max(), got[:, 2].max()) + 1)
    table, max_id = D.read_numeric_literals(files, n, 2)
    want = O.numeric_literal_table([open(f).read() for f in files], max(n, max_id + 1), 2)
    assert np.array_equal(table, want)
