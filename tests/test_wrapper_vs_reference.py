"""The flow-model wrapper (callers of the CUDA path) against the REAL reference,
imported from /root/reference when it is present (build container only; skipped on the GPU box):
identical state-dict layout and seeded init of the full 104.9 M-parameter flow model, and bit-identical
prepared tensors (x1, mask, mu, spks, cond, prompt_lens) handed to compute_loss / the ODE solver for the
same `random` / torch seeds."""
import os
import random
import sys

import pytest
import torch

REF = "/root/reference/cosyvoice_flow_finetune"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def test_wrapper_matches_reference():
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.pop(k) for k in ("flow_model", "utils", "config", "modules", "lora") if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import config as RC
        import flow_model as R
        import utils as RU
        from cosyvoice_lora_finetune_framework_b200 import flow_model as O
        RU.set_all_random_seed(11); a=R.build_flow_model(None,'cpu')
        RU.set_all_random_seed(11); b=O.build_flow_model(None,'cpu')
        sa,sb=a.state_dict(),b.state_dict()
        assert list(sa)==list(sb), set(sa)^set(sb)
        assert all(torch.equal(sa[k],sb[k]) for k in sa)
        print("full flow model: %d keys identical incl. seeded init"%len(sa), sum(p.numel() for p in b.parameters()))
        # capture compute_loss args
        def cap(store):
            def f(x1,mask,mu,spks,cond=None,prompt_lens=None):
                store.update(x1=x1,mask=mask,mu=mu,spks=spks,cond=cond,prompt_lens=prompt_lens); return torch.zeros(()),None
            return f
        g=torch.Generator().manual_seed(0)
        B=3; tl=torch.tensor([40,33,21]); fl=torch.tensor([70,57,36])
        batch=dict(speech_token=torch.randint(0,4096,(B,40),generator=g),speech_token_len=tl,speech_feat=torch.randn(B,70,80,generator=g)*2-6,speech_feat_len=fl,embedding=torch.randn(B,192,generator=g),cross_sample_mel=torch.randn(B,30,80,generator=g)*2-6,cross_sample_mel_len=torch.tensor([30,0,12]))
        a.eval(); b.eval()
        for trial in range(4):
            ra,rb={},{}
            a.decoder.compute_loss=cap(ra); b.decoder.compute_loss=cap(rb)
            random.seed(trial); a(batch,torch.device('cpu')); random.seed(trial); b(batch,torch.device('cpu'))
            assert ra['prompt_lens']==rb['prompt_lens'], (ra['prompt_lens'],rb['prompt_lens'])
            for k in ('x1','mask','mu','spks','cond'): assert torch.equal(ra[k],rb[k]), k
        print("training wrapper: prepared tensors identical over 4 seeds", ra['prompt_lens'])
        # every anti-leakage strategy switch of the per-utterance plan (flow_model.py:301-387): silence gap on, dynamic
        # prompt length off, cross-sample off, text blinding off / always, prompt dropout always
        from cosyvoice_lora_finetune_framework_b200 import config as OC
        variants = [dict(silence_padding_enabled=True), dict(dynamic_prompt_enabled=False),
                    dict(cross_sample_enabled=False, text_blinding_enabled=False),
                    dict(silence_padding_enabled=True, text_blinding_prob=1.0, prompt_max_ratio=0.6),
                    dict(prompt_dropout_prob=1.0)]
        keep_r, keep_o = dict(R.ANTI_LEAKAGE_CONFIG), dict(OC.ANTI_LEAKAGE_CONFIG)
        try:
            for vi, var in enumerate(variants):
                for cfgd, keep in ((R.ANTI_LEAKAGE_CONFIG, keep_r), (OC.ANTI_LEAKAGE_CONFIG, keep_o)):
                    cfgd.clear(); cfgd.update(keep); cfgd.update(var)
                for trial in range(3):
                    ra,rb={},{}
                    a.decoder.compute_loss=cap(ra); b.decoder.compute_loss=cap(rb)
                    random.seed(100+trial); a(batch,torch.device('cpu')); random.seed(100+trial); b(batch,torch.device('cpu'))
                    assert ra['prompt_lens']==rb['prompt_lens'], (var, ra['prompt_lens'],rb['prompt_lens'])
                    for k in ('x1','mask','mu','spks','cond'): assert torch.equal(ra[k],rb[k]), (var, k)
        finally:
            R.ANTI_LEAKAGE_CONFIG.clear(); R.ANTI_LEAKAGE_CONFIG.update(keep_r)
            OC.ANTI_LEAKAGE_CONFIG.clear(); OC.ANTI_LEAKAGE_CONFIG.update(keep_o)
        print("anti-leakage variants identical:", len(variants))
        for mode in ('full','mixed'):
            for mod in (RC,):
                mod.NO_PROMPT_TRAINING_CONFIG.update(enabled=True,mode=mode)
            R.NO_PROMPT_TRAINING_CONFIG.update(enabled=True,mode=mode); O.NO_PROMPT_TRAINING_CONFIG.update(enabled=True,mode=mode)
            ra,rb={},{}; a.decoder.compute_loss=cap(ra); b.decoder.compute_loss=cap(rb)
            random.seed(5); a(batch,torch.device('cpu')); random.seed(5); b(batch,torch.device('cpu'))
            assert ra['prompt_lens']==rb['prompt_lens']
            for k in ('x1','mask','mu','spks','cond'): assert torch.equal(ra[k],rb[k]), k
        print("no-prompt modes identical")
        # inference arg capture
        def capd(store):
            def f(mu,mask,n_timesteps,temperature=1.0,spks=None,cond=None,prompt_len=0,cache=None):
                store.update(mu=mu,mask=mask,n=n_timesteps,spks=spks,cond=cond,prompt_len=prompt_len); return torch.zeros(1,80,mu.shape[2]),torch.zeros(1,80,34,2)
            return f
        ra,rb={},{}
        a.decoder.forward=capd(ra); b.decoder.forward=capd(rb)
        tok=torch.randint(0,4096,(1,120),generator=g); ptok=torch.randint(0,4096,(1,30),generator=g); pf=torch.randn(1,52,80,generator=g); emb=torch.randn(1,192,generator=g)
        a.inference(tok,torch.tensor([120]),ptok,torch.tensor([30]),pf,torch.tensor([52]),emb)
        b.inference(tok,torch.tensor([120]),ptok,torch.tensor([30]),pf,torch.tensor([52]),emb)
        assert ra['n']==rb['n'] and ra['prompt_len']==rb['prompt_len']
        for k in ('mu','mask','spks','cond'): assert torch.allclose(ra[k],rb[k],atol=1e-6), k
        print("inference wrapper identical; n_timesteps", ra['n'], "T", ra['mu'].shape)
        ra.clear(); rb.clear()
        a.inference_like_training(tok,torch.tensor([120]),400,emb,prompt_feat=pf,prompt_len=20)
        b.inference_like_training(tok,torch.tensor([120]),400,emb,prompt_feat=pf,prompt_len=20)
        assert ra['n']==rb['n']==15
        for k in ('mu','mask','spks','cond'): assert torch.allclose(ra[k],rb[k],atol=1e-6), k
        print("inference_like_training identical")
    finally:
        sys.path.remove(REF)
        for k in ("flow_model", "utils", "config", "modules", "lora"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
        from cosyvoice_lora_finetune_framework_b200 import flow_model as O2
        O2.NO_PROMPT_TRAINING_CONFIG.update(enabled=False, mode='full')


def test_lora_inject_and_merge_match_reference_on_the_flow_model():
    """apply_lora_to_model / get_lora_state_dict / get_merged_state_dict on the whole flow model with the reference's
    flow_lora target list: same replaced layers, same seeded LoRA init, same merged keys IN THE SAME ORDER, same values."""
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.pop(k) for k in ("flow_model", "utils", "config", "modules", "lora") if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import flow_model as R
        import lora as RL
        import utils as RU
        from cosyvoice_lora_finetune_framework_b200 import flow_model as O
        from cosyvoice_lora_finetune_framework_b200 import lora as OL
        arch = dict(encoder_num_blocks=2, decoder_n_blocks=1, decoder_num_mid_blocks=2)
        targets = ['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'w_1', 'w_2']
        RU.set_all_random_seed(17)
        a = R.build_flow_model(None, 'cpu', **arch)
        sa = RL.apply_lora_to_model(a, r=16, lora_alpha=32, lora_dropout=0.05, target_modules=targets)
        RU.set_all_random_seed(17)
        b = O.build_flow_model(None, 'cpu', **arch)
        sb = OL.apply_lora_to_model(b, r=16, lora_alpha=32, lora_dropout=0.05, target_modules=targets)
        assert sa == sb, (sa, sb)
        la, lb = RL.get_lora_state_dict(a), OL.get_lora_state_dict(b)
        assert list(la) == list(lb) and all(torch.equal(la[k], lb[k]) for k in la)
        ma, mb = RL.get_merged_state_dict(a), OL.get_merged_state_dict(b)
        assert list(ma.keys()) == list(mb.keys())
        assert all(torch.equal(ma[k], mb[k]) for k in ma)
        # merging mutates the wrapped weights in place in both implementations (reference lora.py:264-279)
        wa = a.decoder.estimator.mid_blocks[0][1][0].attn1.to_q.original_layer.weight
        wb = b.decoder.estimator.mid_blocks[0][1][0].attn1.to_q.original_layer.weight
        assert torch.equal(wa, wb) and torch.equal(wa, ma['decoder.estimator.mid_blocks.0.1.0.attn1.to_q.weight'])
    finally:
        sys.path.remove(REF)
        for k in ("flow_model", "utils", "config", "modules", "lora"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)


def test_public_signatures_match_reference():
    """The drop-in surface of SURVEY section 8b: same parameter names, kinds, order and defaults as the reference
    for every class / function on the path (path-valued defaults excepted: they follow the install location)."""
    import inspect
    sys.dont_write_bytecode = True
    names = ("flow_model", "utils", "config", "modules", "lora", "llm_flow_model", "merge_joint_weights", "dataset")
    saved = {k: sys.modules.pop(k) for k in names if k in sys.modules}
    sys.path.insert(0, REF)
    try:
        import flow_model as RF
        import llm_flow_model as RJ
        import lora as RL
        import merge_joint_weights as RMJ
        import modules as RM
        import utils as RU
        from cosyvoice_lora_finetune_framework_b200 import flow_model as OF
        from cosyvoice_lora_finetune_framework_b200 import llm_flow_model as OJ
        from cosyvoice_lora_finetune_framework_b200 import lora as OL
        from cosyvoice_lora_finetune_framework_b200 import merge_joint_weights as OMJ
        from cosyvoice_lora_finetune_framework_b200 import modules as OM
        from cosyvoice_lora_finetune_framework_b200 import utils as OU
        table = [
            (RF.ConditionalCFM, OF.ConditionalCFM, ["__init__", "compute_loss", "forward", "solve_euler"]),
            (RM.ConditionalDecoder, OM.ConditionalDecoder, ["__init__", "forward"]),
            (RF.MaskedDiffWithXvec, OF.MaskedDiffWithXvec, ["__init__", "forward", "inference", "inference_like_training",
                                                           "normalize_mel", "denormalize_mel"]),
            (RL.LoRALinear, OL.LoRALinear, ["__init__", "forward"]),
            (RJ.JointLLMFlowModel, OJ.JointLLMFlowModel, ["__init__", "forward"]),
            (RF, OF, ["build_flow_model"]),
            (RL, OL, ["apply_lora_to_model", "get_lora_state_dict", "save_lora_weights", "load_lora_weights",
                      "merge_lora_weights", "get_merged_state_dict"]),
            (RU, OU, ["make_pad_mask", "mask_to_bias", "set_all_random_seed"]),
            (RJ, OJ, ["build_joint_model", "get_joint_merged_state_dict"]),
            (RMJ, OMJ, ["find_latest_joint_checkpoint", "merge_flow_from_checkpoint", "merge_llm_from_checkpoint",
                        "merge_both_from_checkpoint"]),
        ]

        def sig(f):
            return [(n, p.kind.name, None if isinstance(p.default, str) and os.sep in p.default else p.default)
                    for n, p in inspect.signature(f).parameters.items()]

        checked = 0
        for ref_owner, our_owner, attrs in table:
            for a in attrs:
                assert sig(getattr(ref_owner, a)) == sig(getattr(our_owner, a)), (getattr(ref_owner, "__name__", ref_owner), a)
                checked += 1
        assert checked == 32
    finally:
        sys.path.remove(REF)
        for k in names:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
