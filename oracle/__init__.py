"""Test infrastructure: CPU oracle for the ProtNote scoring hot path (never imported by protnote_b200)."""
