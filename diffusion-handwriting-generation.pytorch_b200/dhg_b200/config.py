"""Minimal reader for the experiment `config.yml`.

The reference loads it with ruamel.yaml + addict (config.py:36-41) and the
sampling path reads three keys only: training_args.att_layers_num, .channels,
.dropout (checkpoint.py:280-286).  ruamel/addict are not dependencies here; the
file is a two-level block mapping with scalars and flow lists, which this
reader covers.  Missing keys read as None, like the reference's CfgDict.
"""


def _scalar(tok):
    tok = tok.strip()
    if tok == "" or tok in ("~", "null", "Null", "NULL"):
        return None
    if tok in ("true", "True"):
        return True
    if tok in ("false", "False"):
        return False
    if tok.startswith("[") and tok.endswith("]"):
        inner = tok[1:-1].strip()
        return [] if not inner else [_scalar(t) for t in inner.split(",")]
    if (tok[0] == tok[-1]) and tok[0] in "'\"" and len(tok) >= 2:
        return tok[1:-1]
    for cast in (int, float):
        try:
            return cast(tok)
        except ValueError:
            pass
    return tok


def _strip_comment(line):
    out, quote = [], None
    for i, ch in enumerate(line):
        if quote:
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
        elif ch == "#" and (i == 0 or line[i - 1] in " \t"):
            break
        out.append(ch)
    return "".join(out).rstrip()


class CfgDict(dict):
    """dict with attribute access; missing keys -> None."""

    def __getattr__(self, item):
        return self.get(item)

    def __missing__(self, key):
        return None


def parse_yaml(text):
    root = CfgDict()
    stack = [(-1, root)]
    for raw in text.splitlines():
        line = _strip_comment(raw)
        if not line.strip():
            continue
        indent = len(line) - len(line.lstrip(" "))
        key, sep, val = line.strip().partition(":")
        if not sep:
            raise ValueError(f"unsupported YAML line: {raw!r}")
        while stack and indent <= stack[-1][0]:
            stack.pop()
        parent = stack[-1][1]
        if val.strip() == "":
            child = CfgDict()
            parent[key.strip()] = child
            stack.append((indent, child))
        else:
            parent[key.strip()] = _scalar(val)
    # a key with an empty value and no children is null, not {}
    def fix(d):
        for k, v in list(d.items()):
            if isinstance(v, CfgDict):
                if not v:
                    d[k] = None
                else:
                    fix(v)
    fix(root)
    return root


class DLConfig:
    def __init__(self, cfg):
        self._cfg = cfg

    def __getattr__(self, item):
        return getattr(self._cfg, item)

    def __getitem__(self, key):
        return self._cfg[key]

    @classmethod
    def load(cls, path):
        with open(path) as f:
            return cls(parse_yaml(f.read()))

    def update(self, options):
        """Dotted-key overrides, like the reference's cfg_options."""
        for key, value in (options or {}).items():
            node = self._cfg
            parts = key.split(".")
            for p in parts[:-1]:
                if not isinstance(node.get(p), dict):
                    node[p] = CfgDict()
                node = node[p]
            node[parts[-1]] = value
