"""Test stand-in for `matplotlib` (not installed): impl/crossmodal.py:3,43-56 only needs pyplot.subplots /
tight_layout / savefig / close.  savefig writes a small marker file so a test can see it was called.
TEST SCAFFOLDING, not product code."""
