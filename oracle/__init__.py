"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the model(tokens)->logits hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product path (llm_np_cp_b200) never does: it fails loudly when the
CUDA extension is missing.
"""
