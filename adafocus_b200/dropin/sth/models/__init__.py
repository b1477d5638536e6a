"""Drop-in `models` package for the reference's Something-Something tree: put `<repo>/adafocus_b200/dropin/sth` in
front of sys.path and `from models.gfv_net import GFV` in evaluate.py resolves to the B200 implementation; the
reference's `ops/` (dataset, transforms, utils) and `basic_tools/` stay in use."""
