"""Device-resident feature store: the reference's MELData batch assembly
(upstream drin/data.py:85-126) done by index inside the front-end kernel.

CPU: the store's materialised batch equals what the UNMODIFIED reference loader produces from the same cache
directory (skipped when /root/reference is absent).  GPU: the indexed path gives the same bits as running the
materialised batch through the normal 14-tensor entry point."""
import json
import os

import numpy as np
import pytest
import torch

import drin_b200
from drin_b200.store import FeatureStore, synthetic_tables
from oracle import ref_import

FILES = {  # store table name -> reference cache file (drin/data.py:44-70,158-200)
    "mention_text_feature": "mention-text-feature_{s}.npy", "mention_start_pos": "start-pos_{s}.npy",
    "mention_end_pos": "end-pos_{s}.npy", "mention_image_feature": "mention-image-feature_{s}.npy",
    "mention_object_feature": "mention-object-feature_{s}.npy", "mention_object_score": "mention-object-score_{s}.npy",
    "miet_similarity": "similarity-miet_{s}.npy", "mtei_similarity": "similarity-eimt_{s}.npy", "answer": "answer_{s}.npy",
}
WD_ENTITY = {"entity_text_feature": "entity-attr-feature_{s}.npy", "entity_image_feature": "entity-image-feature_{s}.npy",
             "entity_object_feature": "entity-object-feature_{s}.npy", "entity_object_score": "entity-object-score_{s}.npy"}
WM_ENTITY = {"entity_text_feature": "entity-attr-feature.npy", "entity_text_mask": "entity-attr-mask.npy",
             "entity_image_feature": "entity-image-feature_all.npy", "entity_object_feature": "entity-object-feature_all.npy",
             "entity_object_score": "entity-object-score_all.npy"}


def write_cache_dir(path, dataset, tables, split, C):
    """Write synthetic tables under the reference's file names (flat entity files for WikiDiverse, like
    preprocess/*.py emit them; a qid -> row json for WikiMEL)."""
    os.makedirs(path, exist_ok=True)
    for k, f in FILES.items():
        np.save(os.path.join(path, f.format(s=split)), tables[k].numpy())
    np.save(os.path.join(path, f"mention-text-mask_{split}.npy"),
            np.ones(tables["mention_text_feature"].shape[:2], dtype=np.int64))
    if dataset == "wikidiverse":
        for k, f in WD_ENTITY.items():
            t = tables[k]
            np.save(os.path.join(path, f.format(s=split)), t.reshape((-1,) + tuple(t.shape[2:])).numpy())   # [N*C, ...]
    else:
        for k, f in WM_ENTITY.items():
            np.save(os.path.join(path, f), tables[k].numpy())
        Ne = tables["entity_text_feature"].shape[0]
        perm = torch.randperm(Ne, generator=torch.Generator().manual_seed(5))
        qid = [f"Q{int(1000 + i)}" for i in range(Ne)]
        with open(os.path.join(path, "qid2idx.json"), "w") as fh:
            json.dump({qid[i]: int(i) for i in perm.tolist()}, fh)
        names = np.array([[qid[j] for j in row] for row in tables["entity_index"].tolist()])
        np.save(os.path.join(path, f"entity-name-raw_{split}.npy"), names.reshape(-1))


def make_tables(dataset):
    if dataset == "wikidiverse":
        return synthetic_tables(dataset, 9, 3, 10), 11
    return synthetic_tables(dataset, 7, 4, 5, num_entities=20, entity_tokens=16, mention_tokens=32), 6


@pytest.mark.parametrize("dataset", ["wikidiverse", "wikimel"])
def test_from_preprocess_dir_round_trip(tmp_path, dataset):
    tables, C = make_tables(dataset)
    write_cache_dir(str(tmp_path), dataset, tables, "train", C)
    a = FeatureStore(dataset, tables, C, device="cpu")
    b = FeatureStore.from_preprocess_dir(str(tmp_path), "train", dataset, C, device="cpu")
    idx = [4, 0, 6, 6]
    for x, y in zip(a.batch(idx), b.batch(idx)):
        assert x.dtype == y.dtype and torch.equal(x, y)
    assert len(a) == tables["answer"].numel() and a.nbytes() > 0
    with pytest.raises(IndexError):
        a.select([len(a)])


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("dataset", ["wikidiverse", "wikimel"])
def test_store_batch_equals_reference_loader(tmp_path, dataset):
    """Same cache directory through the reference's MELData + default collate and through the store."""
    import sys

    tables, C = make_tables(dataset)
    write_cache_dir(str(tmp_path), dataset, tables, "valid", C)
    le = tables["entity_text_feature"].shape[1] if dataset == "wikimel" else 64
    ref_import.load(dataset, C - 1, le, preprocess_dir=str(tmp_path), mention_mmap=None, entity_mmap=None,
                    max_mention_sentence_len=tables["mention_text_feature"].shape[1], dataloader_workers=0,
                    batch_size=4, shuffle_train_data=False)
    import drin.data as ref_data
    from torch.utils.data import default_collate

    loaders = None
    try:
        onehot = np.concatenate([np.eye(C - 1, dtype=np.uint8), np.zeros((1, C - 1), dtype=np.uint8)], 0)   # data.py:159-161
        if dataset == "wikimel":
            ent = [np.load(os.path.join(tmp_path, WM_ENTITY[k])) for k in
                   ("entity_text_feature", "entity_text_mask", "entity_image_feature", "entity_object_feature", "entity_object_score")]
        else:
            ent = [np.load(os.path.join(tmp_path, WD_ENTITY["entity_text_feature"].format(s="valid"))), None] + [
                np.load(os.path.join(tmp_path, WD_ENTITY[k].format(s="valid")))
                for k in ("entity_image_feature", "entity_object_feature", "entity_object_score")]
        ds = ref_data.MELData((onehot, ent[0], ent[1], ent[2], ent[3], ent[4], "valid"))
        idx = [5, 2, 2, 0]
        want = default_collate([ds[i] for i in idx])
    finally:
        for name in [m for m in sys.modules if m.split(".")[0] in ("common", "baselines", "drin")]:
            del sys.modules[name]
    got = FeatureStore.from_preprocess_dir(str(tmp_path), "valid", dataset, C, device="cpu").batch(idx)
    assert len(got) == len(want) == 15
    for i, (g, w) in enumerate(zip(got, want)):
        if i == 8 and dataset == "wikidiverse":
            assert tuple(g.shape) == tuple(w.shape) and int(w.abs().sum()) == 0     # collated int 0
            continue
        assert tuple(g.shape) == tuple(w.shape), (i, g.shape, w.shape)
        assert torch.equal(g, w.to(g.dtype)), i


# ------------------------------------------------------------------------------------------------
# GPU: indexed gather == materialised batch, bit for bit
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dataset,dtype", [("wikidiverse", torch.float32), ("wikimel", torch.float32),
                                           ("wikidiverse", torch.bfloat16), ("wikimel", torch.bfloat16)])
def test_indexed_path_is_bit_identical_to_materialised_batch(dataset, dtype):
    tables, C = make_tables(dataset)
    store = FeatureStore(dataset, tables, C, device="cuda", feature_dtype=dtype)
    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=C).cuda()
    tr = drin_b200.Trainer(model)
    idx = [3, 0, 5, 5, 1]
    s_idx = tr.rank_scores(store.select(idx)).clone()
    s_mat = tr.rank_scores(store.batch(idx)).clone()
    assert torch.equal(s_idx, s_mat)
    l_idx = tr.forward_backward(store.select(idx)).clone()
    g_idx = model.flat_grads.clone()
    l_mat = tr.forward_backward(store.batch(idx)).clone()
    assert torch.equal(l_idx, l_mat) and torch.equal(g_idx, model.flat_grads)
    assert float(g_idx.abs().max()) > 0


@pytest.mark.gpu
def test_indexed_training_matches_oracle():
    from oracle import drin_oracle as O
    from drin_b200.synthetic import spread_weights

    tables, C = make_tables("wikidiverse")
    store = FeatureStore("wikidiverse", tables, C, device="cuda")
    cpu = FeatureStore("wikidiverse", tables, C, device="cpu")
    cfg = O.DrinConfig(num_candidates_model=C)
    sd = spread_weights(O.init_state(cfg, 0))
    model = drin_b200.Model(num_candidates_model=C)
    model.load_state_dict(sd)
    tr = drin_b200.Trainer(model.cuda(), margin=cfg.triplet_margin)
    idx = [8, 1, 4, 2, 7, 0]
    loss = tr.forward_backward(store.select(idx))
    b = cpu.batch(idx)
    s_ref, l_ref, g_ref = O.train_step_grads(sd, b[:-1], b[-1], cfg)
    assert abs(float(loss) - float(l_ref)) <= 1e-4 * abs(float(l_ref))
    assert float((tr.last_scores.cpu() - s_ref).abs().max() / s_ref.abs().max()) < 1e-4
    for k, g in model._grad_views().items():
        assert float((g.cpu() - g_ref[k]).abs().max() / g_ref[k].abs().max()) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("edge_feature", ["scaler", "vector"])
def test_graphed_step_replays_bit_identically_to_eager_steps(edge_feature):
    """The captured CUDA graph of the whole step (front end .. Adam, device-side step count) gives the same bits as
    eager steps over the same index batches (scalar and vector edges)."""
    tables, C = make_tables("wikidiverse")
    store = FeatureStore("wikidiverse", tables, C, device="cuda")
    batches = [[0, 3, 5, 8], [1, 2, 6, 7], [4, 4, 0, 2]]
    outs = []
    for graphed in (False, True):
        torch.manual_seed(0)
        model = drin_b200.Model(num_candidates_model=C, gcn_edge_feature=edge_feature).cuda()
        tr = drin_b200.Trainer(model, lr=1e-3)
        losses = []
        if graphed:
            gs = drin_b200.GraphedStoreStep(tr, store, 4)
            for b in batches:
                losses.append(float(gs.step(b)))
            assert tr.opt.step_count == 3
        else:
            for b in batches:
                losses.append(float(tr.step(store.select(b))))
        outs.append((losses, model.flat_params.clone()))
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1])
