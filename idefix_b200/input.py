"""idefix.ini reader with the reference's grammar and accessors.

Mirrors ``Input`` (src/input.cpp:30-110 grammar, src/input.hpp:15-69 accessors):
``[Block]`` headers, ``name v0 v1 ...`` entries, ``#`` comments anywhere.
"""
from ._abi import IdefixError


class Input:
    def __init__(self, filename=None, text=None):
        self.inputFileName = filename or "idefix.ini"
        self.inputParameters = {}
        if text is None:
            try:
                with open(self.inputFileName) as fh:
                    text = fh.read()
            except OSError:
                raise IdefixError(f"Input: cannot open input file {self.inputFileName}")
        self._parse(text)

    @classmethod
    def from_string(cls, text):
        return cls(filename="<string>", text=text)

    def _parse(self, text):
        block = None
        nparams = 0
        for raw in text.splitlines():
            line = raw.split("#", 1)[0]
            if not line.strip(" "):
                continue
            s = line.lstrip(" ")
            if s.startswith("["):
                end = s.find("]", 1)
                if end < 0:
                    raise IdefixError(f"Block name '{block}' in file '{self.inputFileName}' not properly ended")
                if block is not None and nparams == 0:
                    self.inputParameters[block]["!!empty!!"] = ["!!empty!!"]
                block = s[1:end]
                self.inputParameters.setdefault(block, {})
                nparams = 0
                continue
            if block is None:
                raise IdefixError(
                    f"Input file '{self.inputFileName}' must specify a block name before the first parameter")
            tok = line.split()
            nparams += 1
            self.inputParameters[block].setdefault(tok[0], []).extend(tok[1:])

    # -- accessors (src/input.hpp:26-69) ------------------------------------------
    def CheckBlock(self, block):
        return block in self.inputParameters

    def CheckEntry(self, block, name):
        """number of values of the entry, -1 when absent (src/input.cpp:316-330)"""
        try:
            return len(self.inputParameters[block][name])
        except KeyError:
            return -1

    def Get(self, typ, block, name, num):
        if self.CheckEntry(block, name) <= num:
            raise IdefixError(f"Mandatory parameter [{block}]:{name}({num}) cannot be found in the input file.")
        raw = self.inputParameters[block][name][num]
        try:
            if typ is bool:
                low = raw.lower()
                if low in ("yes", "true", "1"):
                    return True
                if low in ("no", "false", "0"):
                    return False
                raise ValueError(raw)
            if typ is int:
                return int(float(raw)) if any(c in raw for c in ".eE") else int(raw)
            return typ(raw)
        except ValueError:
            raise IdefixError(f"Cannot convert [{block}]:{name}({num}) = '{raw}' to {typ.__name__}")

    def GetOrSet(self, typ, block, name, num, default):
        """src/input.hpp:147-170.  The default is stored through an ostream (6 significant digits)
        and read back, so e.g. the default gamma 5/3 becomes 1.66667 -- reproduced here."""
        size = self.CheckEntry(block, name)
        if size <= num:
            if size < 0 and num > 0:
                raise IdefixError(f"Entry [{block}]:{name}is not defined.\n"
                                  "Only the first (index 0) parameter can be set by default.")
            if size + 1 < num:
                raise IdefixError(f"Entry [{block}]:{name}has {size} parameters.")
            if isinstance(default, bool):
                text = "true" if default else "false"
            elif isinstance(default, float):
                text = "%g" % default
            else:
                text = str(default)
            self.inputParameters.setdefault(block, {}).setdefault(name, []).append(text)
        return self.Get(typ, block, name, num)

    def set(self, block, name, *values):
        """replace an entry (what a test driver does when it edits the ini file)"""
        self.inputParameters.setdefault(block, {})[name] = [str(v) for v in values]

    def remove(self, block, name):
        self.inputParameters.get(block, {}).pop(name, None)

    def dumps(self):
        out = []
        for block, params in self.inputParameters.items():
            out.append(f"[{block}]")
            for name, vals in params.items():
                if name != "!!empty!!":
                    out.append(f"{name}    " + "  ".join(vals))
            out.append("")
        return "\n".join(out)
