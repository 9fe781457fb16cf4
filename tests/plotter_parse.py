"""Test helpers: benchmark/src/bin/plotter.rs's log parser restated, and an fd-level stderr capture."""
import os
import re
import tempfile

def capture_fd2(fn):
    """Runs fn() with file descriptor 2 redirected to a file; returns (result, text)."""
    import sys

    sys.stderr.flush()
    saved = os.dup(2)
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        os.dup2(tmp.fileno(), 2)
        try:
            res = fn()
        finally:
            os.dup2(saved, 2)
            os.close(saved)
        tmp.seek(0)
        return res, tmp.read().decode("utf-8")


def plotter_parse(text):
    """Log::parse of benchmark/src/bin/plotter.rs:337-373 restated: returns the top-level logs as
    [(name, depth, duration_ns, [children...])]."""
    def parse_duration(d):  # plotter.rs:504-515
        m = re.search(r"\d(?=\D*$)", d)
        value, unit = d[: m.end()], d[m.end():]
        return float(value) * {"ns": 1.0, "µs": 1e3, "ms": 1e6, "s": 1e9}[unit]

    stack, logs = [], []
    for line in text.splitlines():
        indent, sep, log = line.rpartition("·")
        if not sep:
            indent, log = "", line
        if len(log) < 9:
            continue
        prefix, log = log[:9], log[9:]
        depth = (len(indent.encode("utf-8")) + 2) // 4
        if depth == len(stack) and prefix.startswith("Start:"):
            stack.append({"name": log, "depth": depth, "ns": 0.0, "children": []})
        elif prefix.startswith("End:"):
            name = log.rsplit(" ", 1)[0]
            dur = parse_duration(log.rsplit("..", 1)[1]) if ".." in log else 0.0
            idx = max((i for i, l in enumerate(stack) if l["name"] == name), default=None)
            if idx is not None:
                del stack[idx + 1:]
                done = stack.pop()
                done["ns"] = dur
                (stack[-1]["children"] if stack else logs).append(done)
    return logs


