"""Writes tests/golden/cli_flags.json: every sub-command flag of the reference's command line (`faster_qwen3_tts/cli.py`,
build_parser) with its default and whether it is required.  Run in the build container (the reference does not travel):

    python tests/golden/make_cli_golden.py

soundfile is not in this image; the reference's cli imports it at module level, so an empty stand-in module is registered
first (nothing of it is called while the parser is built)."""
import argparse
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def flag_table(parser: argparse.ArgumentParser, prefix: str = "") -> dict:
    out = {}
    for a in parser._actions:
        if isinstance(a, argparse._SubParsersAction):
            for name, sp in a.choices.items():
                out.update(flag_table(sp, prefix + name + " "))
        elif not isinstance(a, argparse._HelpAction):
            for o in a.option_strings:
                out[prefix + o] = {"default": a.default, "required": bool(a.required), "action": type(a).__name__,
                                   "choices": list(a.choices) if a.choices else None}
    return out


if __name__ == "__main__":
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    sys.path.insert(0, "/root/reference")  # the cli imports its own package
    spec = importlib.util.spec_from_file_location("reference_cli", "/root/reference/faster_qwen3_tts/cli.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    table = flag_table(mod.build_parser())
    with open(os.path.join(HERE, "cli_flags.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(len(table), "flags")
