"""Flow half of the reference's no-prompt inference entry point (inference_joint.py:63-234): load a
merged `flow_merged_*.pt` strictly into the stock-layout flow model and run token -> mel with the
mel (de)normalisation patch the reference applies around `flow.inference` (:129-150). The text
front-end, the LLM and the HiFT vocoder of the full pipeline are out of scope (SURVEY section 2 rows
15-18); this module therefore starts from speech tokens and stops at the mel.

  python -m cosyvoice_lora_finetune_framework_b200.inference_joint --flow flow_merged_joint.pt --tokens tok.pt --output mel.pt
"""
import argparse
import os

import torch

from .config import MEL_MEAN, MEL_STD, OUTPUT_DIR, PRETRAINED_MODEL_DIR
from .flow_model import build_flow_model


def load_merged_flow(flow_path=None, device='cuda', dtype=torch.float16):
    model = build_flow_model(PRETRAINED_MODEL_DIR if os.path.isdir(PRETRAINED_MODEL_DIR) else None, device='cpu')
    if flow_path:
        model.load_state_dict(torch.load(flow_path, map_location='cpu'), strict=True)   # inference_joint.py:124
    model = model.to(device).eval()
    model.decoder.estimator.cvflow_dtype = dtype
    return model


@torch.inference_mode()
def flow_inference_normalized(model, token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len,
                              embedding, flow_cache=None):
    """`flow.inference` with prompt-mel normalisation before and de-normalisation after (:129-150)."""
    if prompt_feat is not None and prompt_feat.shape[1] > 0:
        prompt_feat = (prompt_feat - MEL_MEAN) / MEL_STD
    mel, cache = model.inference(token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len,
                                 embedding, flow_cache)
    return mel * MEL_STD + MEL_MEAN, cache


def inference_no_prompt_joint(speech_token, flow_path=None, embedding=None, device='cuda', n_timesteps=10):
    """speech tokens (1, N) -> log-mel (1, 80, T) with zero speaker embedding and empty prompts,
    the call the reference makes through `model.tts(..., zero embedding, empty prompts)` (:191-201)."""
    model = load_merged_flow(flow_path, device)
    dev = next(model.parameters()).device
    token = speech_token.to(dev)
    emb = torch.zeros(1, 192, device=dev) if embedding is None else embedding.to(dev)
    empty_tok = torch.zeros(1, 0, dtype=torch.int32, device=dev)
    empty_feat = torch.zeros(1, 0, 80, device=dev)
    mel, _ = flow_inference_normalized(model, token, torch.tensor([token.shape[1]], device=dev), empty_tok,
                                       torch.tensor([0], device=dev), empty_feat, torch.tensor([0], device=dev), emb)
    return mel


def main():
    ap = argparse.ArgumentParser(description='flow half of the no-prompt joint inference')
    ap.add_argument('--flow', type=str, default=os.path.join(OUTPUT_DIR, 'flow_merged_joint.pt'))
    ap.add_argument('--tokens', type=str, required=True, help='torch file holding a (1, N) speech-token tensor')
    ap.add_argument('--output', type=str, default='mel.pt')
    a = ap.parse_args()
    mel = inference_no_prompt_joint(torch.load(a.tokens), a.flow if os.path.exists(a.flow) else None)
    torch.save(mel.cpu(), a.output)
    print("mel", tuple(mel.shape), "->", a.output)


if __name__ == '__main__':
    main()
