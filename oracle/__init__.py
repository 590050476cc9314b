"""Test-infrastructure oracle for the NISQA predict hot path (see nisqa_oracle.py header).

Never imported by the product package ``nisqa_b200``."""
