"""ORACLE: CPU restatement of the reference hot path.  Test infrastructure only (see frontend_ref.py)."""
