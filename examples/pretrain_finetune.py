#!/usr/bin/env python
"""BASELINE.json configs[1]: a Pet_v2-shaped KG with synthetic literal features of the reference dims, pre-training
(TransR loss) + fine-tuning (BPR loss) + link-prediction evaluation on one B200 -- the reference's main.py loop
(main.py:80-317) written against the drop-in classes, with the minibatches assembled on the device.

    python examples/pretrain_finetune.py [--entities 770554 --edges 1050000 --relations 15 --layers 8 --steps 20]
    python examples/pretrain_finetune.py --preset small     # BASELINE.json configs[0]: data/Small shape, ONE
                                                            # pre-training epoch (triples / batch steps) + evaluation
    python examples/pretrain_finetune.py --data-dir /root/reference/data/Test   # a reference data directory, where present

The Pet_v2 train-split KG blob and the Drive-hosted literal pickles are not available offline, so the graph and the
literal tables are generated with the shapes of SURVEY.md 8(d) cfg 2 (power-law heads capped at out-degree 50).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import literalkg_b200 as L


def reference_args(a):
    """argument.py defaults that the model reads."""
    return argparse.Namespace(
        use_pretrain=0, device="cuda", embed_dim=300, relation_dim=300, scale_gat_dim=256, use_residual=True,
        alpha=0.1, lamda=0.5, aggregation_type=a.aggregator, n_conv_layers=a.layers, conv_dim=32, mess_dropout=0.1,
        kg_l2loss_lambda=1e-5, fine_tuning_l2loss_lambda=1e-5, pre_training_neg_rate=3, fine_tuning_neg_rate=3,
        num_lit_dim=2, txt_lit_dim=300, use_num_lit=True, use_txt_lit=True, milestone_score=0.5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entities", type=int, default=770_554)
    ap.add_argument("--edges", type=int, default=1_050_000)
    ap.add_argument("--relations", type=int, default=15)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--aggregator", default="bi-interaction")
    ap.add_argument("--steps", type=int, default=20, help="optimizer steps per phase")
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--diseases", type=int, default=118, help="candidate tails of the fine-tuning task")
    ap.add_argument("--seed", type=int, default=2022)
    ap.add_argument("--ft-steps", type=int, default=None, help="fine-tuning steps (default: --steps)")
    ap.add_argument("--preset", choices=["pet_v2", "small"], default="pet_v2",
                    help="small: the shape of the reference's bundled data/Small split (770 563 entity ids, 251 895 triples, "
                         "15 relations, out-degree <= 49, 118 disease tails) and main.py's epoch length")
    ap.add_argument("--data-dir", default=None, help="read a reference data directory (dataloader.py:24-32 layout) "
                                                     "instead of generating the KG; literals missing there are synthetic")
    ap.add_argument("--kg-file", default="pre_training_train.txt")
    a = ap.parse_args()
    if a.preset == "small":
        a.entities, a.edges, a.relations = 770_563, 251_895, 15
        a.steps = a.edges // a.batch + 1                   # main.py:107: n_kg_batch = n_kg_train // batch_size + 1
        a.ft_steps = 0 if a.ft_steps is None else a.ft_steps
    a.ft_steps = a.steps if a.ft_steps is None else a.ft_steps
    dev = torch.device("cuda:0")
    torch.manual_seed(a.seed)
    if a.data_dir:
        trip, num_table, text_table = L.dataloader.read_data_dir(a.data_dir, a.kg_file)
        kg = argparse.Namespace(h=trip[:, 0], r=trip[:, 1], t=trip[:, 2])
        a.entities, a.relations = int(max(kg.h.max(), kg.t.max())) + 1, int(kg.r.max()) + 1
        a.entities = max([a.entities] + [tab.shape[0] for tab in (num_table, text_table) if tab is not None])
    else:
        kg = L.synthetic.make_kg(a.entities, a.edges, a.relations, seed=a.seed, max_out_degree=50)
    n = a.entities
    num, txt = L.synthetic.make_literals(n, seed=a.seed, device=dev)
    data = L.KGTensors(kg.h, kg.t, kg.r, n_entities=n, device=dev)
    if a.data_dir and num_table is not None:
        num = data._table(num_table)                       # the bundled numeric literal dictionaries, min-max scaled
    args = reference_args(a)
    model = L.LiteralKG(args, n, a.relations, data.A_in, num, txt).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    # fine-tuning pairs: animals (heads of the KG) x a small set of disease entities
    rng = np.random.default_rng(a.seed)
    diseases = rng.choice(n, a.diseases, replace=False)
    ft_heads = np.unique(kg.h)[: 20_000]
    ft_pairs = np.unique(np.stack([np.repeat(ft_heads, 2), rng.choice(diseases, 2 * len(ft_heads))], 1), axis=0)
    ft_plan = L.GraphPlan(torch.from_numpy(ft_pairs[:, 0]).to(dev), torch.from_numpy(ft_pairs[:, 1]).to(dev),
                          torch.zeros(len(ft_pairs), dtype=torch.int64, device=dev), n, 1)
    kg_sampler = L.BatchSampler(data.plan, np.unique(kg.t), args.pre_training_neg_rate, True, seed=a.seed)
    ft_sampler = L.BatchSampler(ft_plan, diseases, args.fine_tuning_neg_rate, False, seed=a.seed)

    def sync():
        torch.cuda.synchronize(dev)

    def phase(name, step_fn, steps):
        if steps <= 0:
            return []
        model.train()
        losses = []
        step_fn(); sync()                                  # warm-up (plan / cache construction)
        t0 = time.perf_counter()
        for _ in range(steps):
            losses.append(step_fn())
        sync()
        dt = (time.perf_counter() - t0) / steps
        losses = [float(x) for x in losses]
        print(f"{name}: {dt * 1e3:.1f} ms / step (sampling + forward + backward + Adam), "
              f"loss {losses[0]:.4f} -> {losses[-1]:.4f}")
        return losses

    def pre_step():                                        # main.py:112-124
        h, r, p, ng = kg_sampler.sample(a.batch)
        opt.zero_grad(set_to_none=True)
        loss = model(h, r, p, ng, device=dev, mode="pre_training")
        loss.backward()
        opt.step()
        return loss.detach()

    def ft_step():                                         # main.py:213-226
        h, _, p, ng = ft_sampler.sample(a.batch)
        opt.zero_grad(set_to_none=True)
        loss = model(h, p, ng, device=dev, mode="fine_tuning")
        loss.backward()
        opt.step()
        return loss.detach()

    t0 = time.perf_counter()
    with torch.no_grad():                                  # main.py:147-151, once per epoch
        model(data.h_list, data.t_list, data.r_list, data.relations, device=dev, mode="update_att")
    sync()
    print(f"update_att: {(time.perf_counter() - t0) * 1e3:.1f} ms (includes the plan build)")
    pre = phase("pre_training", pre_step, a.steps)
    ft = phase("fine_tuning", ft_step, a.ft_steps)

    # evaluation (utils/model_utils.py:41-78): every head batch calls mode='predict'; the embedding pass behind it is
    # computed once and reused while the parameters do not change
    model.eval()
    heads = torch.from_numpy(ft_heads[:4096]).to(dev)
    tails = torch.from_numpy(np.sort(diseases)).to(dev)
    with torch.no_grad():
        sync(); t0 = time.perf_counter()
        preds = [model(heads[i:i + 2048], tails, device=dev, mode="predict") for i in range(0, len(heads), 2048)]
        vals, pos, _ = model.topk(heads, tails, 10)
        sync()
    print(f"evaluation: {(time.perf_counter() - t0) * 1e3:.1f} ms for {len(heads)} heads x {len(tails)} tails "
          f"(predict + top-10), positive rate {torch.cat(preds).float().mean().item():.3f}")
    assert np.isfinite(pre + ft).all() and samplers_ok(kg_sampler, ft_sampler)
    print("ok")


def samplers_ok(*samplers):
    return all(int(s.n_failed.item()) == 0 for s in samplers)


if __name__ == "__main__":
    main()
