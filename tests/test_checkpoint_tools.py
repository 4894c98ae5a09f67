"""Checkpoint tooling of the path, CPU only: the merge tool turns a Lightning-layout joint
checkpoint (`state_dict` keys `model.flow.*`, reference train_joint.py:312-320) into a stock-layout `flow_merged.pt`
(reference merge_joint_weights.py:122-176) that loads strictly into an un-wrapped flow model."""
import functools
import os
import time

import torch

from cosyvoice_lora_finetune_framework_b200 import flow_model, llm_flow_model, lora, merge_joint_weights
from cosyvoice_lora_finetune_framework_b200.config import JOINT_TRAINING_CONFIG

SMALL = dict(encoder_num_blocks=1, decoder_n_blocks=1, decoder_num_mid_blocks=1)


def test_merge_flow_from_checkpoint_layout_and_values(tmp_path, monkeypatch):
    small_build = functools.partial(flow_model.build_flow_model, **SMALL)
    monkeypatch.setattr(llm_flow_model, "build_flow_model", small_build)
    cfg = JOINT_TRAINING_CONFIG["flow_lora"]
    torch.manual_seed(3)
    trained = llm_flow_model.build_joint_model(None, "cpu", "flow_only", None, cfg)
    with torch.no_grad():                                   # "training": move the LoRA factors
        for n, p in trained.named_parameters():
            if "lora_" in n:
                p.add_(0.05 * torch.randn_like(p))
    sd = trained.state_dict()
    assert any(k.startswith("flow.encoder.") and k.endswith("lora_A") for k in sd)       # encoder LoRA targets too
    ckpt = str(tmp_path / "joint_flow_only_last.ckpt")
    torch.save({"state_dict": {"model." + k: v.clone() for k, v in sd.items()}, "epoch": 0}, ckpt)
    out = str(tmp_path / "flow_merged.pt")
    torch.manual_seed(99)                                    # the tool rebuilds the model: its random init must not matter
    merged = merge_joint_weights.merge_flow_from_checkpoint(ckpt, out)
    stock = small_build(None, "cpu")
    assert set(merged.keys()) == set(stock.state_dict().keys())
    assert not any("lora_" in k or "original_layer" in k for k in merged)
    s = cfg["lora_alpha"] / cfg["lora_r"]
    for prefix in ("decoder.estimator.mid_blocks.0.1.0.attn1.to_k", "encoder.encoders.0.self_attn.linear_q",
                   "encoder.encoders.0.feed_forward.w_1"):
        w0 = sd["flow.%s.original_layer.weight" % prefix]
        want = w0 + (sd["flow.%s.lora_B" % prefix] @ sd["flow.%s.lora_A" % prefix]) * s
        assert torch.allclose(merged[prefix + ".weight"], want, atol=1e-6), prefix
    # frozen tensors pass through untouched; the file round-trips and loads strictly into the stock layout
    assert torch.equal(merged["decoder.estimator.final_proj.weight"], sd["flow.decoder.estimator.final_proj.weight"])
    stock.load_state_dict(torch.load(out, map_location="cpu"), strict=True)
    # get_joint_merged_state_dict: only the halves that carry LoRA (reference llm_flow_model.py:313-336)
    both = llm_flow_model.get_joint_merged_state_dict(trained)
    assert set(both) == {"flow"} and set(both["flow"].keys()) == set(merged.keys())


def test_find_latest_joint_checkpoint(tmp_path):
    d = str(tmp_path)
    for i, name in enumerate(["joint_flow_only_last.ckpt", "joint_joint_03_0.4100.ckpt", "joint_llm_only_last.ckpt", "notes.txt"]):
        p = os.path.join(d, name)
        open(p, "w").close()
        os.utime(p, (time.time() + i, time.time() + i))
    f = merge_joint_weights.find_latest_joint_checkpoint
    assert os.path.basename(f(d)) == "joint_joint_03_0.4100.ckpt"                 # joint runs preferred when no mode given
    assert os.path.basename(f(d, "flow_only")) == "joint_flow_only_last.ckpt"
    assert os.path.basename(f(d, "llm_only")) == "joint_llm_only_last.ckpt"
    os.remove(os.path.join(d, "joint_joint_03_0.4100.ckpt"))
    assert os.path.basename(f(d)) == "joint_llm_only_last.ckpt"                   # else simply the newest
    assert f(str(tmp_path / "..") if False else d, "nothing_like_this") is None


def test_lora_adapter_roundtrip_through_flow_model(tmp_path):
    """save_lora_weights / load_lora_weights (reference lora.py:239-256) on the whole flow model."""
    a = flow_model.build_flow_model(None, "cpu", **SMALL)
    b = flow_model.build_flow_model(None, "cpu", **SMALL)
    for m in (a, b):
        lora.apply_lora_to_model(m, r=4, lora_alpha=8, lora_dropout=0.0, target_modules=["to_q", "to_v", "linear_k", "w_2"])
    path = str(tmp_path / "adapter.pt")
    lora.save_lora_weights(a, path)
    lora.load_lora_weights(b, path)
    sa, sb = lora.get_lora_state_dict(a), lora.get_lora_state_dict(b)
    assert list(sa) == list(sb) and len(sa) > 0 and all(torch.equal(sa[k], sb[k]) for k in sa)
