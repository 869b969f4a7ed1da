"""Global config dict, same keys / defaults as the reference's gpd.py:4-133 for the hot path.

The architecture sizes are constants of the compiled kernels (the reference also freezes them at
import time); `temperature`, `max_len`, `verbose`, `use_cuda` are honoured at
call time like the reference does (model.py:834, 819; main.py:43,123)."""

gpd = {
    'verbose': True,
    # audio (gpd.py:7-21)
    'sample_rate': 16000,
    'bit_depth': 16,
    'window_len': .025,
    'window_step': .01,
    'n_mels': 80,
    'preemphasis': .97,
    'delta_delta': True,
    'downsample': True,
    'normalize': True,
    # dictionary (gpd.py:38-46)
    'pad': 0, 'sos': 1, 'eos': 2, 'unk': 3,
    'max_num_words': 5000,
    # encoder / decoder / attention (gpd.py:56-93) - fixed by the kernels
    'encoder_type': 'LSTM',
    'skip_step': 0,
    'encoder_hidden_size': 256,
    'encoder_num_layers': 4,
    'residual': True,
    'encoder_bidirectional': True,
    'decoder_type': 'LSTM',
    'decoder_hidden_size': 512,
    'decoder_num_layers': 1,
    'embed_dim': 256,
    'temperature': 1.,
    'input_feeding': True,
    'dec_init_cell_state_as_param': False,
    'attn_type': 'B',
    'attn_size': 128,
    'map_enc': False,
    'heads': 1,
    # eval (gpd.py:113-127)
    'beam_width': 4,
    'second_pass': True,
    'max_len': 40,
    'lm_weight': 0.0,
    'length_weight': 0.0,
    # added at run time by main.py:123-124
    'use_cuda': True,
    'eval_num_workers': 0,
    'eval_batch_size': 256,      # gpd.py:118
}

_FROZEN = {
    'sample_rate': 16000, 'window_len': .025, 'window_step': .01, 'n_mels': 80,
    'delta_delta': True, 'downsample': True, 'encoder_type': 'LSTM', 'skip_step': 0,
    'encoder_hidden_size': 256, 'encoder_num_layers': 4, 'residual': True,
    'encoder_bidirectional': True, 'decoder_type': 'LSTM', 'decoder_hidden_size': 512,
    'decoder_num_layers': 1, 'embed_dim': 256, 'input_feeding': True, 'attn_type': 'B',
    'attn_size': 128, 'map_enc': False, 'heads': 1, 'max_num_words': 5000,
    'pad': 0, 'sos': 1, 'eos': 2, 'unk': 3,
}


def check_frozen():
    """The kernels implement exactly the default architecture; refuse anything else loudly."""
    bad = {k: gpd[k] for k, v in _FROZEN.items() if gpd.get(k) != v}
    if bad:
        raise ValueError(f"asr_b200 kernels are compiled for the reference defaults; unsupported gpd overrides: {bad}")
