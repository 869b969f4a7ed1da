"""B200-native (sm_100a) inference hot path of shawnthu/chinese-asr.

Mirrors the reference's module namespaces for the path named in BASELINE.json:
    gpd    - the global config dict                       (reference gpd.py)
    data   - AudioBase, MelScale, get_log_mel             (reference data.py:21-280, 371-382)
             AudioDst, AudioLoader (eval / infer), read_pcm (reference data.py:109-121, 385-540)
    model  - Model.load / eval_one_batch_with_greedy/_beam (reference model.py:18-82, 357-369, 503-987)
             Model.wer, Model.test_model                  (reference util.py:237-262, model.py:1370-1439)
    parallel - shard_utterances / gather_hypotheses (one process per GPU), BatchPipeline (engines per GPU)
    main   - ASR(lm_path, bw), parse                      (reference main.py:27-102)
    lm     - NGramLM: second-pass LM with device tables    (replaces kenlm at model.py:755)
Every computation runs in hand-written CUDA kernels behind the C ABI of include/asr_b200.h
(libasr_b200.so); there is no CPU or PyTorch fallback - importing without the library raises.
"""
from . import _cabi  # noqa: F401  (fails loudly when libasr_b200.so is missing)
from .gpd import gpd  # noqa: F401
