"""Minimal stand-in for timm==0.3.2 (pinned by Painter/requirements.txt:1, asserted at
Painter/main_train.py:24).  TEST INFRASTRUCTURE ONLY: lets the unmodified reference import here."""
__version__ = "0.3.2"
