"""TEST INFRASTRUCTURE ONLY: run pytest in THIS process with every emulated test module loading the sanitized build of
the kernels (tests/emu/build_emu.py, sanitize=True).  Started by tests/test_emu_sanitize.py with libasan preloaded:

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/emu/san_runner.py <pytest args>
    D3H_EMU_VARIANT=tsan LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python tests/emu/san_runner.py <pytest args>
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
import pytest  # noqa: E402

_variant = os.environ.get("D3H_EMU_VARIANT", "san")      # "san": ASan + UBSan; "tsan": ThreadSanitizer fibres
_san = build_emu.build(sanitize=_variant == "san", tsan=_variant == "tsan")
build_emu.build = lambda force=False, sanitize=False, tsan=False: _san
sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider"] + sys.argv[1:]))
