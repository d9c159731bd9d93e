"""TEST INFRASTRUCTURE ONLY: CPU oracle for the spVIPES training hot path (see restatement.py).
Never imported by the product package spvipes_b200/."""
