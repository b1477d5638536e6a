"""Drop-in `models` package for the reference's ACT tree: put `<repo>/adafocus_b200/dropin/act` in front of
sys.path (before the reference's own directory) and `from models.gfv_net import GFV` in main_dist.py resolves to the
B200 implementation; `ops/` and `basic_tools/` remain the reference's."""
