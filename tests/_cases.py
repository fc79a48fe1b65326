"""Seeded parity cases shared by the oracle-pinning tests and the GPU parity tests."""
import numpy as np

from svdfeature_b200 import synth

NU, NI, NG = 200, 100, 30
BASE = dict(num_user=NU, num_item=NI, num_factor=16, learning_rate=0.01, wd_user=0.004, wd_item=0.004,
            base_score=3.6)


def _binary(d):
    return (d[0], (d[1] > 3).astype(np.float32), d[2], d[3])


def as_ugroup(csr, rows_per_block=10):
    n = len(csr[1])
    bro = np.arange(0, n + 1, rows_per_block, dtype=np.int32)
    if bro[-1] != n:
        bro = np.append(bro, np.int32(n))
    nb = len(bro) - 1
    return (bro, np.zeros(nb + 1, np.int32), np.zeros(nb, np.int32), np.zeros(0, np.uint32),
            np.zeros(0, np.float32)) + tuple(csr)


def split_tags(ug, every=3):
    """Re-tag a user-grouped batch so that some users span START/MIDDLE/END blocks
    (apex_svd_data.h:353-371): every `every`-th block is cut in up to three pieces that
    all carry the user's feedback list."""
    bro, bfo, tag, fi, fv = ug[:5]
    nbro, nbfo, ntag, nfi, nfv = [0], [0], [], [], []
    for b in range(len(tag)):
        r0, r1 = int(bro[b]), int(bro[b + 1])
        f = slice(int(bfo[b]), int(bfo[b + 1]))
        cuts = [r0, r1]
        if b % every == 0 and r1 - r0 >= 3:
            third = (r1 - r0) // 3
            cuts = [r0, r0 + third, r0 + 2 * third, r1]
        pieces = len(cuts) - 1
        for p in range(pieces):
            nbro.append(cuts[p + 1])
            nfi.append(fi[f]); nfv.append(fv[f])
            nbfo.append(nbfo[-1] + (f.stop - f.start))
            ntag.append(0 if pieces == 1 else (1 if p == 0 else (2 if p == pieces - 1 else 3)))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return (np.asarray(nbro, np.int32), np.asarray(nbfo, np.int32), np.asarray(ntag, np.int32),
            cat(nfi, np.uint32), cat(nfv, np.float32)) + tuple(ug[5:])


def cases():
    """name -> (format_type, active_type, params, data, kind['csr'|'ug'])"""
    out = {}
    out["basic_k16"] = (0, 0, BASE, synth.basic_mf(5000, NU, NI, seed=1), "csr")
    out["basic_k64"] = (0, 0, dict(BASE, num_factor=64), synth.basic_mf(5000, NU, NI, seed=11), "csr")
    out["basic_k128"] = (0, 0, dict(BASE, num_factor=128), synth.basic_mf(2000, NU, NI, seed=12), "csr")
    out["basic_k256"] = (0, 0, dict(BASE, num_factor=256), synth.basic_mf(1500, NU, NI, seed=13), "csr")
    gen = dict(BASE, num_factor=13, wd_user_bias=0.01, wd_item_bias=0.02, num_global=NG, wd_global=0.001)
    out["general_k13_dups"] = (0, 0, gen, synth.random_general(3000, NU, NI, NG, seed=2, allow_dup=True), "csr")
    out["general_k40"] = (0, 0, dict(gen, num_factor=40, num_regfree_global=7),
                          synth.random_general(3000, NU, NI, NG, seed=3), "csr")
    out["many_features"] = (0, 0, dict(gen, num_factor=24),
                            synth.random_general(600, NU, NI, NG, seed=4, max_g=25, max_u=40, max_i=45,
                                                 allow_dup=True), "csr")
    out["no_decay"] = (0, 0, dict(BASE, wd_user=0.0, wd_item=0.0), synth.basic_mf(3000, NU, NI, seed=5), "csr")
    out["no_user_bias"] = (0, 0, dict(gen, no_user_bias=1), synth.random_general(2000, NU, NI, NG, seed=6), "csr")
    for act in (1, 2, 3, 5, 6, 7):
        pa = dict(gen, base_score=0.6, num_factor=8)
        out["active_%d" % act] = (0, act, pa, _binary(synth.random_general(2000, NU, NI, NG, seed=30 + act)), "csr")
    out["neighborhood_k32"] = (0, 0, dict(gen, num_factor=32),
                               synth.neighborhood(3000, NU, NI, NG, ng=8, seed=7), "csr")
    pw = dict(BASE, no_user_bias=1, base_score=0.5, num_factor=32)
    out["pairwise_csr"] = (0, 3, pw, synth.pairwise(4000, NU, NI, seed=8), "csr")
    out["pairwise_ugroup"] = (1, 3, pw, as_ugroup(synth.pairwise(4000, NU, NI, seed=8)), "ug")
    pp = dict(BASE, num_ufeedback=NI, wd_ufeedback=0.004, wd_ufeedback_bias=0.001, scale_lr_ufeedback=0.5,
              ufeedback_init_sigma=0.01)
    out["svdpp_k16"] = (1, 0, pp, synth.user_grouped(4000, NU, NI, avg_fb=15, seed=9), "ug")
    out["svdpp_k64_tags"] = (1, 0, dict(pp, num_factor=64),
                             split_tags(synth.user_grouped(3000, NU, NI, avg_fb=40, seed=10)), "ug")
    out["svdpp_no_user_bias"] = (1, 0, dict(pp, no_user_bias=1), synth.user_grouped(2000, NU, NI, avg_fb=10, seed=14), "ug")
    out["lr_decay"] = (0, 0, dict(BASE, decay_learning_rate=1, decay_rate=0.9), synth.basic_mf(3000, NU, NI, seed=15), "csr")
    # regularisers other than L2 decay (base.h:188-283): L1 soft threshold, projection, L1-user/L2-item,
    # L1 on the global bias, non-negative user factors
    out["reg_l1"] = (0, 0, dict(gen, reg_method=1, reg_global=1, wd_user=0.02, wd_item=0.03, wd_global=0.05),
                     synth.random_general(2500, NU, NI, NG, seed=16, allow_dup=True), "csr")
    out["reg_project"] = (0, 0, dict(gen, num_factor=24, reg_method=2, wd_user=0.0015, wd_item=0.002),
                          synth.random_general(2500, NU, NI, NG, seed=17), "csr")
    out["reg_l1_user_only"] = (0, 0, dict(gen, num_factor=20, reg_method=3, wd_user=0.02),
                               synth.random_general(2500, NU, NI, NG, seed=18, allow_dup=True), "csr")
    out["reg_project_basic_k64"] = (0, 0, dict(BASE, num_factor=64, reg_method=2, wd_user=0.005, wd_item=0.006),
                                    synth.basic_mf(4000, NU, NI, seed=19), "csr")
    out["user_nonnegative"] = (0, 0, dict(gen, num_factor=16, user_nonnegative=1, item_nonnegative=1),
                               synth.random_general(2500, NU, NI, NG, seed=20), "csr")
    out["svdpp_reg_l1"] = (1, 0, dict(pp, reg_method=1, wd_user=0.02, wd_item=0.02),
                           synth.user_grouped(2500, NU, NI, avg_fb=12, seed=21), "ug")
    # ranged weight decay (ParameterSet, base.h:33-75): user / item / global index ranges with their own wd
    rng_pairs = [("up:wd", 0.02), ("up:bound", 50), ("up:wd", 0.0), ("up:bound", 120), ("up:wd", 0.008),
                 ("up:bound", NU), ("ip:wd", 0.001), ("ip:bound", 10), ("ip:wd", 0.03), ("ip:bound", NI),
                 ("gp:wd", 0.05), ("gp:bound", 12), ("gp:wd", 0.0005), ("gp:bound", NG)]
    out["ranged_wd"] = (0, 0, dict(gen, num_factor=20, num_regfree_global=3, wd_ranges=rng_pairs),
                        synth.random_general(3000, NU, NI, NG, seed=22, allow_dup=True), "csr")
    out["ranged_wd_l1"] = (0, 0, dict(gen, reg_method=1, reg_global=1, wd_ranges=rng_pairs),
                           synth.random_general(2500, NU, NI, NG, seed=23), "csr")
    out["ranged_wd_project_uip"] = (0, 0, dict(BASE, num_factor=64, reg_method=2,
                                               wd_ranges=[("uip:wd", 0.002), ("uip:bound", 60), ("uip:wd", 0.004),
                                                          ("uip:bound", NU)]),
                                    synth.basic_mf(4000, NU, NI, seed=24), "csr")
    out["svdpp_ranged_wd"] = (1, 0, dict(pp, wd_ranges=rng_pairs[:10]),
                              synth.user_grouped(2500, NU, NI, avg_fb=12, seed=25), "ug")
    return out


def write_side_features(path, num_row, num_target, seed, max_extra=3, covered=0.7):
    """A feature_user / feature_item file (apex-utils/apex_utils.h:176-195): line r lists the extra
    (index:value) pairs feature index r expands to; rows beyond the file have none."""
    rng = np.random.default_rng(seed)
    rows = int(num_row * covered)
    side = []
    with open(path, "w") as f:
        for r in range(rows):
            n = int(rng.integers(0, max_extra + 1))
            idx = rng.integers(0, num_target, n)
            val = np.round(rng.uniform(-1.0, 1.0, n), 3)
            side.append((idx.astype(np.uint32), val.astype(np.float32)))
            f.write("%d %s\n" % (n, " ".join("%d:%g" % (i, v) for i, v in zip(idx, val))))
    return side


# cases whose arithmetic calls expf: the GPU is allowed 1 ulp of expf (see svdgpu.h)
SIGMOID_CASES = {"active_1", "active_2", "active_3", "active_7", "pairwise_csr", "pairwise_ugroup"}


def hparams_of(params, base_score):
    keys = ("learning_rate", "wd_user", "wd_item", "wd_user_bias", "wd_item_bias", "wd_global", "reg_method",
            "reg_global", "num_regfree_global", "scale_lr_ufeedback", "wd_ufeedback", "wd_ufeedback_bias",
            "user_nonnegative")
    hp = {k: params[k] for k in keys if k in params}
    hp["base_score"] = base_score
    return hp


def shape_of(params, fmt, act):
    return dict(num_user=params["num_user"], num_item=params["num_item"], num_factor=params["num_factor"],
                num_global=params.get("num_global", 0), num_ufeedback=params.get("num_ufeedback", 0),
                no_user_bias=params.get("no_user_bias", 0), active_type=act, format_type=fmt)


def rank_cases():
    """name -> (format_type, model params, stream kwargs, ranker params) for SVDFeatureRanker parity.
    The model is trained for one round on a small batch first so that no tensor is at its initial value."""
    gen = dict(BASE, num_factor=20, num_global=NG, wd_global=0.001)
    pp = dict(BASE, num_factor=16, num_ufeedback=NI, wd_ufeedback=0.004, ufeedback_init_sigma=0.05)
    out = {}
    out["rank_pos_k20"] = (0, gen, dict(num_item_set=70, num_sections=25, num_global=NG, seed=1), {})
    out["rank_top5_k20"] = (0, gen, dict(num_item_set=70, num_sections=25, num_global=NG, seed=2), {"top_k": 5})
    out["rank_top3_k64"] = (0, dict(BASE, num_factor=64), dict(num_item_set=300, num_sections=40, seed=3), {"top_k": 3})
    out["rank_pos_k13"] = (0, dict(gen, num_factor=13), dict(num_item_set=129, num_sections=30, num_global=NG, seed=4,
                                                             max_pos=9, max_ban=20), {})
    out["rank_svdpp_pos"] = (1, pp, dict(num_item_set=60, num_sections=20, seed=5, ugroup=True, num_ufeedback=NI), {})
    out["rank_svdpp_top4_split"] = (1, pp, dict(num_item_set=60, num_sections=20, seed=6, ugroup=True,
                                                num_ufeedback=NI, split_every=2), {"top_k": 4})
    return out


def rank_model(fmt, params, tmpdir, cls=None):
    """Train a model of the case's shape with the oracle and save it: the file every ranker loads."""
    import os
    from _oracle import COracle
    o = COracle(fmt, 0, 0, params)
    o.init(7)
    nu, ni, ng = params["num_user"], params["num_item"], params.get("num_global", 0)
    if fmt == 1:
        o.update_ugroup(synth.user_grouped(1500, nu, ni, avg_fb=8, seed=31))
    else:
        o.update_csr(synth.random_general(1500, nu, ni, ng, seed=31) if ng else synth.basic_mf(1500, nu, ni, seed=31))
    path = os.path.join(str(tmpdir), "rank.model")
    o.save_model(path)
    return path
