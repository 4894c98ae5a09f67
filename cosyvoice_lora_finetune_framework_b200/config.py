"""Configuration dictionaries of the flow / joint LoRA fine-tune (same names and default values as
the reference's config.py, which its modules import by name). Values the hot path reads:
ANTI_LEAKAGE_CONFIG['boundary_*'] (flow_model.compute_loss) and MEL_MEAN / MEL_STD."""
import os

PROJECT_ROOT = os.path.dirname(os.path.abspath(__file__))


def _first_existing(candidates, probe):
    for c in candidates:
        if os.path.exists(os.path.join(c, probe)):
            return c
    return candidates[0]


PRETRAINED_MODEL_DIR = _first_existing(
    [os.path.join(PROJECT_ROOT, "pretrained_models", "CosyVoice-300M"),
     os.path.join(os.path.dirname(PROJECT_ROOT), "pretrained_models", "CosyVoice-300M")], "flow.pt")
DATA_DIR = os.path.join(PROJECT_ROOT, "data")
RAW_AUDIO_DIR = os.path.join(PROJECT_ROOT, "raw_audio")
OUTPUT_DIR = os.path.join(PROJECT_ROOT, "output")

TRAIN_CONFIG = {
    'max_epochs': 100, 'batch_size': 2, 'accumulate_grad_batches': 4, 'learning_rate': 1e-4,
    'min_learning_rate': 1e-6, 'weight_decay': 0.01, 'warmup_steps': 50, 'max_feat_len': 600,
    'precision': '16-mixed', 'gradient_clip_val': 1.0, 'augmentation': True,
}

LORA_CONFIG = {
    'use_lora': True, 'lora_r': 16, 'lora_alpha': 16, 'lora_dropout': 0.05,
    'target_modules': ['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'linear_out', 'w_1', 'w_2'],
}

ANTI_LEAKAGE_CONFIG = {
    'silence_padding_enabled': False, 'silence_token_id': 0, 'silence_min_tokens': 5, 'silence_max_tokens': 10,
    'silence_mel_value': -11.5,
    'dynamic_prompt_enabled': True, 'prompt_min_ratio': 0.05, 'prompt_max_ratio': 0.20,
    'prompt_dropout_enabled': True, 'prompt_dropout_prob': 0.25,
    'boundary_loss_enabled': True, 'boundary_frames': 25, 'boundary_loss_weight': 5.0,
    'cross_sample_enabled': True, 'cross_sample_prob': 0.85,
    'text_blinding_enabled': True, 'text_blinding_prob': 0.95, 'text_blinding_mode': 'zero',
}

NO_PROMPT_TRAINING_CONFIG = {'enabled': False, 'mode': 'full', 'no_prompt_ratio': 0.8, 'use_mean_embedding': False}

JOINT_TRAINING_CONFIG = {
    'training_mode': 'joint', 'llm_loss_weight': 2.0, 'flow_loss_weight': 1.0, 'no_prompt_training': True,
    'llm_lora': {'lora_r': 8, 'lora_alpha': 16, 'lora_dropout': 0.15,
                 'target_modules': ['linear_q', 'linear_k', 'linear_v', 'linear_out', 'w_1', 'w_2']},
    'flow_lora': {'lora_r': 16, 'lora_alpha': 32, 'lora_dropout': 0.05,
                  'target_modules': ['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'w_1', 'w_2']},
    'learning_rate': 2e-4, 'max_epochs': 100, 'batch_size': 1, 'accumulate_grad_batches': 16, 'max_feat_len': 250,
}

MEL_MEAN = -6.0
MEL_STD = 2.0

INFERENCE_CONFIG = {
    'max_prompt_seconds': 5, 'physical_trim_enabled': True, 'physical_trim_mode': 'absolute',
    'physical_trim_frames': 80, 'physical_trim_extra_ms': 300, 'trim_ratio': 0.08, 'boundary_trim_ratio': 0.20,
}

MODEL_CONFIG = {'input_size': 512, 'output_size': 80, 'spk_embed_dim': 192, 'vocab_size': 4096,
                'input_frame_rate': 50, 'sample_rate': 22050}
