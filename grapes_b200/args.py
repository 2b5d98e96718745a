"""``Arguments`` of the reference (/root/reference/main.py:23-54) and its ``--config_file`` handling
(main.py:367-373), without the ``typed-argument-parser`` dependency (absent here, no network).

Behaviour kept from Tap 1.8.0 with ``explicit_bool=True`` as the reference uses it:
  * every field is ``--name value`` or ``--name=value``; booleans take an explicit value that is a
    case-insensitive prefix of "true"/"false" or "1"/"0" (configs use ``true``, ``True``, ``false``);
  * a config file is a whitespace/shlex-split argument string (quoted values such as ``--dataset "cora"``
    occur, configs/gflownet/cora.txt:2) that is PREPENDED to the command line, so CLI flags win;
  * unknown flags are an error.
"""
from __future__ import annotations

import dataclasses
import shlex
import sys
from typing import List, Optional, Sequence


def _parse_bool(s: str) -> bool:
    v = s.strip().lower()
    if v and ("true".startswith(v) or v == "1"):
        return True
    if v and ("false".startswith(v) or v == "0"):
        return False
    raise ValueError(f"invalid boolean value {s!r}")


@dataclasses.dataclass
class Arguments:
    dataset: str = 'cora'

    sampling_hops: int = 2
    num_samples: int = 16
    use_indicators: bool = True
    lr_gf: float = 1e-4
    lr_gc: float = 1e-3
    loss_coef: float = 1e4
    log_z_init: float = 0.
    reg_param: float = 0.
    dropout: float = 0.

    model_type: str = 'gcn'
    hidden_dim: int = 256
    embed_nodes: bool = False
    node_emb_dim: int = 64
    max_epochs: int = 30
    batch_size: int = 512
    eval_frequency: int = 5
    eval_on_cpu: bool = True
    eval_full_batch: bool = True
    random_sampling: bool = False

    runs: int = 10
    split_id: int = 0
    seed: Optional[int] = None
    notes: Optional[str] = None
    log_wandb: bool = False
    config_file: Optional[str] = None

    reinforce_baseline: bool = False

    # ------------------------------------------------------------------
    def as_dict(self) -> dict:
        return dataclasses.asdict(self)

    @classmethod
    def _types(cls):
        out = {}
        for f in dataclasses.fields(cls):
            t = f.type if isinstance(f.type, str) else getattr(f.type, "__name__", str(f.type))
            out[f.name] = t
        return out

    @classmethod
    def _convert(cls, name: str, raw: str):
        t = cls._types()[name]
        if "bool" in t:
            return _parse_bool(raw)
        if "int" in t:
            return int(raw)
        if "float" in t:
            return float(raw)
        return raw

    @classmethod
    def _apply(cls, obj: "Arguments", argv: Sequence[str]):
        names = cls._types()
        i = 0
        argv = list(argv)
        while i < len(argv):
            tok = argv[i]
            if not tok.startswith("--"):
                raise SystemExit(f"error: unrecognized arguments: {tok}")
            if "=" in tok:
                key, raw = tok[2:].split("=", 1)
                i += 1
            else:
                key = tok[2:]
                if i + 1 >= len(argv):
                    raise SystemExit(f"error: argument --{key}: expected one argument")
                raw = argv[i + 1]
                i += 2
            if key not in names:
                raise SystemExit(f"error: unrecognized arguments: --{key}")
            try:
                setattr(obj, key, cls._convert(key, raw))
            except ValueError as e:
                raise SystemExit(f"error: argument --{key}: {e}")

    @classmethod
    def parse_args(cls, argv: Optional[Sequence[str]] = None, config_files: Optional[List[str]] = None) -> "Arguments":
        argv = list(sys.argv[1:] if argv is None else argv)
        file_args: List[str] = []
        for path in config_files or []:
            with open(path) as f:
                file_args += shlex.split(f.read(), comments=True)
        obj = cls()
        cls._apply(obj, file_args + argv)
        return obj


def parse_cli(argv: Optional[Sequence[str]] = None) -> Arguments:
    """main.py:367-373: parse once; if ``--config_file`` is given, parse again with the file prepended so
    the command line takes precedence."""
    args = Arguments.parse_args(argv)
    if args.config_file is not None:
        args = Arguments.parse_args(argv, config_files=[args.config_file])
    return args
