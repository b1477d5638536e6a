"""Minimal stand-in for `omegaconf` (see oracle/stubs/hydra): basic_tools/__init__.py:7 imports these two names."""
import yaml


class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def to_yaml(args):
        d = dict(vars(args)) if not isinstance(args, dict) else dict(args)
        return yaml.safe_dump({k: (v if isinstance(v, (int, float, str, bool, type(None), list)) else str(v))
                               for k, v in d.items()})
