"""Joint LLM + flow fine-tuning glue (API mirror of the reference's llm_flow_model.py).

The flow branch (`_forward_flow`, reference :181-229) is the caller of the CUDA hot path: it prepares
(feat, mask, mu, spk, zero cond) and calls `flow.decoder.compute_loss`. The LLM branch (:109-179) is
host orchestration over an opaque `llm` module (TransformerLM of upstream CosyVoice, out of scope
here) and is restated only so that `training_mode='joint'|'llm_only'` keep
working when such a module is supplied.
"""
from typing import Any, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import JOINT_TRAINING_CONFIG, MEL_MEAN, MEL_STD, PRETRAINED_MODEL_DIR
from .flow_model import build_flow_model
from .lora import apply_lora_to_model, get_merged_state_dict
from .utils import make_pad_mask

IGNORE_ID = -1   # cosyvoice.utils.common.IGNORE_ID


def th_accuracy(pad_outputs: torch.Tensor, pad_targets: torch.Tensor, ignore_label: int) -> torch.Tensor:
    """Token accuracy over non-ignored targets (upstream cosyvoice/utils/common.py)."""
    pred = pad_outputs.view(pad_targets.size(0), pad_targets.size(1), pad_outputs.size(1)).argmax(2)
    keep = pad_targets != ignore_label
    hit = torch.sum(pred.masked_select(keep) == pad_targets.masked_select(keep))
    return (hit / torch.sum(keep)).detach()


class JointLLMFlowModel(nn.Module):
    """training_mode: 'joint' | 'llm_only' | 'flow_only'; returns {'loss', 'llm_loss', 'flow_loss', 'llm_acc'}."""

    def __init__(self, llm: Optional[nn.Module], flow: nn.Module, training_mode: str = 'joint',
                 llm_loss_weight: float = 1.0, flow_loss_weight: float = 1.0, no_prompt_training: bool = True):
        super().__init__()
        self.llm = llm
        self.flow = flow
        self.training_mode = training_mode
        self.llm_loss_weight = llm_loss_weight
        self.flow_loss_weight = flow_loss_weight
        self.no_prompt_training = no_prompt_training
        self.mel_mean = MEL_MEAN
        self.mel_std = MEL_STD

    def normalize_mel(self, mel: torch.Tensor) -> torch.Tensor:
        return (mel - self.mel_mean) / self.mel_std

    def forward(self, batch: dict, device: torch.device) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        if self.training_mode in ('joint', 'llm_only'):
            r = self._forward_llm(batch, device)
            out['llm_loss'] = r['loss'] * self.llm_loss_weight
            if 'acc' in r:
                out['llm_acc'] = r['acc']
        if self.training_mode in ('joint', 'flow_only'):
            out['flow_loss'] = self._forward_flow(batch, device)['loss'] * self.flow_loss_weight
        if self.training_mode == 'joint':
            out['loss'] = out['llm_loss'] + out['flow_loss']
        else:
            out['loss'] = out['llm_loss'] if self.training_mode == 'llm_only' else out['flow_loss']
        return out

    def _forward_llm(self, batch: dict, device: torch.device) -> Dict[str, Any]:
        """No-prompt LLM step: [SOS, spk, text..., TASK, speech...] -> next speech token (+EOS)."""
        if self.llm is None:
            raise RuntimeError("training_mode=%r needs an LLM module" % self.training_mode)
        llm = self.llm
        text, text_len = batch['text_token'].to(device), batch['text_token_len'].to(device)
        speech, speech_len = batch['speech_token'].to(device), batch['speech_token_len'].to(device)
        dtype = next(llm.parameters()).dtype
        targets = [torch.tensor([IGNORE_ID] * (2 + int(text_len[i])) + speech[i, :speech_len[i]].tolist() +
                                [llm.speech_token_size]) for i in range(text.size(0))]
        lm_target = torch.nn.utils.rnn.pad_sequence(targets, batch_first=True, padding_value=IGNORE_ID).to(device)
        text_emb, text_emb_len = llm.encode(llm.text_embedding(text), text_len)
        spk = llm.spk_embed_affine_layer(F.normalize(batch['embedding'].to(device).to(dtype), dim=1)).unsqueeze(1)
        sos = llm.llm_embedding.weight[llm.sos_eos].reshape(1, 1, -1)
        task = llm.llm_embedding.weight[llm.task_id].reshape(1, 1, -1)
        lm_in, lm_in_len = llm.pad_unpad_sequence(sos, spk, text_emb, text_emb_len, task, llm.speech_embedding(speech),
                                                  speech_len)
        hidden, _ = llm.llm(lm_in, lm_in_len.to(device))
        logits = llm.llm_decoder(hidden)
        return {'loss': llm.criterion_ce(logits, lm_target),
                'acc': th_accuracy(logits.view(-1, llm.speech_token_size + 1), lm_target, ignore_label=IGNORE_ID)}

    def _forward_flow(self, batch: dict, device: torch.device) -> Dict[str, Any]:
        """No-prompt flow step (reference :181-229): zero conditioning, loss over all valid frames."""
        flow = self.flow
        dtype = flow.input_embedding.weight.dtype
        token, token_len = batch['speech_token'].to(device), batch['speech_token_len'].to(device)
        feat = self.normalize_mel(batch['speech_feat'].to(device).to(dtype))
        feat_len = batch['speech_feat_len'].to(device)
        spk = flow.spk_embed_affine_layer(F.normalize(batch['embedding'].to(device).to(dtype), dim=1))
        keep = (~make_pad_mask(token_len)).to(dtype).unsqueeze(-1).to(device)
        h, _ = flow.encoder(flow.input_embedding(torch.clamp(token, min=0)) * keep, token_len)
        h, _ = flow.length_regulator(flow.encoder_proj(h), feat_len)
        conds = torch.zeros(feat.shape, device=device, dtype=dtype).transpose(1, 2)
        loss_mask = (~make_pad_mask(feat_len)).to(h)
        loss, _ = flow.decoder.compute_loss(feat.transpose(1, 2).contiguous(), loss_mask.unsqueeze(1),
                                            h.transpose(1, 2).contiguous(), spk, cond=conds)
        return {'loss': loss}


def build_joint_model(pretrained_path: str = PRETRAINED_MODEL_DIR, device: str = 'cuda', training_mode: str = 'joint',
                      llm_lora_config: Optional[dict] = None, flow_lora_config: Optional[dict] = None
                      ) -> JointLLMFlowModel:
    """Load llm + flow, inject LoRA per config, wrap (reference :232-310). The LLM comes from the
    upstream `cosyvoice` package when it is importable; `flow_only` needs only this package."""
    llm = None
    if training_mode in ('joint', 'llm_only'):
        try:
            from cosyvoice.cli.cosyvoice import CosyVoice   # upstream pipeline object, not part of this repo
        except ImportError as e:
            raise RuntimeError("training_mode=%r needs the upstream `cosyvoice` package for the LLM (%s); "
                               "use training_mode='flow_only' for the flow path alone" % (training_mode, e))
        cv = CosyVoice(pretrained_path, load_jit=False, load_trt=False)
        llm = cv.model.llm
        if llm_lora_config:
            apply_lora_to_model(llm, r=llm_lora_config.get('lora_r', 8), lora_alpha=llm_lora_config.get('lora_alpha', 16),
                                lora_dropout=llm_lora_config.get('lora_dropout', 0.05),
                                target_modules=llm_lora_config.get('target_modules', ['linear_q', 'linear_k', 'linear_v',
                                                                                     'linear_out', 'w_1', 'w_2']))
    flow = build_flow_model(pretrained_path if pretrained_path and __import__('os').path.exists(pretrained_path) else None,
                            device='cpu')
    if training_mode in ('joint', 'flow_only') and flow_lora_config:
        stats = apply_lora_to_model(flow, r=flow_lora_config.get('lora_r', 16),
                                    lora_alpha=flow_lora_config.get('lora_alpha', 16),
                                    lora_dropout=flow_lora_config.get('lora_dropout', 0.05),
                                    target_modules=flow_lora_config.get('target_modules', ['to_q', 'to_k', 'to_v',
                                                                                          'linear_q', 'linear_k',
                                                                                          'linear_v', 'linear_out',
                                                                                          'w_1', 'w_2']))
        print(f"  Flow LoRA: {stats['replaced_layers']} layers, {stats['trainable_params']:,} params "
              f"({stats['trainable_ratio']:.2f}%)")
    cfg = JOINT_TRAINING_CONFIG
    model = JointLLMFlowModel(llm=llm, flow=flow, training_mode=training_mode,
                              llm_loss_weight=cfg.get('llm_loss_weight', 1.0),
                              flow_loss_weight=cfg.get('flow_loss_weight', 1.0),
                              no_prompt_training=cfg.get('no_prompt_training', True))
    return model.to(device)


def get_joint_merged_state_dict(model: JointLLMFlowModel) -> Dict[str, dict]:
    """{'llm': merged_sd, 'flow': merged_sd} for whichever halves carry LoRA (reference :313-336)."""
    out = {}
    if model.llm is not None and any('lora_' in n for n, _ in model.llm.named_parameters()):
        out['llm'] = get_merged_state_dict(model.llm)
    if any('lora_' in n for n, _ in model.flow.named_parameters()):
        out['flow'] = get_merged_state_dict(model.flow)
    return out
