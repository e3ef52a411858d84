"""Reference import path ``forward_utils`` (test_last.py:17-22, train.py:17-22) -> aa-clip-iqm_b200/forward_utils.py."""
import model  # noqa: F401  (puts the repository root on sys.path and registers aaclip_b200)
from aaclip_b200.forward_utils import (  # noqa: F401
    CLASS_NAMES, DOMAINS, PROMPTS, REAL_NAMES, anomaly_maps, calculate_seg_loss, calculate_similarity_map,
    get_adapted_single_class_text_embedding, get_adapted_single_sentence_text_embedding, get_adapted_text_embedding,
    metrics_eval, visualize)
from aaclip_b200.tokenizer import tokenize  # noqa: F401
