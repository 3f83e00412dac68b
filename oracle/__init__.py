"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's binarized forward path.

Importable from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs only; nothing under ``binary-networks-pytorch_b200/`` may import this package.
"""
