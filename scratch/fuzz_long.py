"""Long differential-fuzz campaign on the CPU: the repo's own property tests (device code through tests/emu == oracle ==
independent python implementations) with far more examples and fresh seeds than the test suite spends.

    python scratch/fuzz_long.py --minutes 30 --jobs 8 [--only test_bind] [--asan]

Every job takes tests round-robin; hypothesis tests get max_examples raised and a per-round seed, seed-parametrized tests get
seeds beyond the suite's.  A failure prints the test, the seed and the traceback and is appended to scratch/fuzz_failures.log."""
import argparse
import importlib
import multiprocessing as mp
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HYP = [  # (module, test, examples per round)
    ("tests.test_emu_parity", "test_random_strings_property", 1500),
    ("tests.test_emu_parity", "test_random_queries_property", 1500),
    ("tests.test_emu_parity", "test_random_paths_property", 1500),
    ("tests.test_route", "test_random_tables_property", 1500),
    ("tests.test_bind", "test_emu_random_bodies", 2000),
    ("tests.test_bind", "test_emu_valid_json_roundtrip", 1000),
    ("tests.test_bind", "test_syntax_verdict_agrees_with_python_json", 2000),
    ("tests.test_bind", "test_emu_bind_float_against_python_float", 1500),
    ("tests.test_http_parse", "test_accepted_messages_against_h11_and_llhttp", 1000),
    ("tests.test_http_parse", "test_mutated_messages_device_code_equals_oracle", 2000),
    ("tests.test_proto", "test_random_message_types_three_way", 1000),
    ("tests.test_proto", "test_decode_random_wire_three_way", 1500),
    ("tests.test_reqlog", "test_emu_random_records_property", 1500),
    ("tests.test_result", "test_string_outcome_property", 1000),
    ("tests.test_grpc", "test_emu_random_messages", 1500),
    ("tests.test_slots", "test_emu_slots_random_rows_property", 1500),
    ("tests.test_slots", "test_emu_slots_random_requests_property", 1500),
    ("tests.test_slots", "test_emu_slots_mixed_stream_property", 400),
]
SEEDED = [  # (module, test, seeds per round)
    ("tests.test_proto_nested", "test_random_types_three_ways", 6),
    ("tests.test_proto_nested", "test_decode_mutated_frames_device_code_equals_oracle", 6),
    ("tests.test_values", "test_random_schemas_three_ways", 3),
    ("tests.test_values", "test_mutated_rows_device_code_equals_oracle", 3),
]


def log_failure(text):
    with open(os.path.join(ROOT, "scratch", "fuzz_failures.log"), "a") as f:
        f.write(text + "\n")
    print(text, flush=True)


def worker(job, jobs, deadline, only, counts):
    from hypothesis import settings, seed as hseed, HealthCheck
    items = [("h",) + t for t in HYP] + [("s",) + t for t in SEEDED]
    if only:
        items = [t for t in items if any(o in t[1] + "." + t[2] for o in only)]
    rnd = 0
    while time.time() < deadline:
        for k, (kind, mod, name, amount) in enumerate(items):
            if k % jobs != (job + rnd) % jobs or time.time() >= deadline:
                continue
            m = importlib.import_module(mod)
            fn = getattr(m, name)
            base = 1_000_003 * (rnd + 1) + 7919 * job + k
            with open(f"/tmp/fuzz_job{job}.now", "w") as nf:   # what was running, should a sanitizer abort the process
                nf.write(f"{mod}.{name} base={base}\n")
            try:
                if kind == "h":
                    inner = fn
                    inner._hypothesis_internal_use_settings = settings(max_examples=amount, deadline=None, database=None,
                                                                      suppress_health_check=list(HealthCheck), derandomize=False)
                    inner._hypothesis_internal_use_seed = base
                    inner()
                    counts[mod + "." + name] = counts.get(mod + "." + name, 0) + amount
                else:
                    for s in range(amount):
                        sd = 100_000 + base * 16 + s
                        try:
                            fn(sd)
                        except Exception:
                            log_failure(f"FAIL {mod}.{name} seed={sd}\n{traceback.format_exc()}")
                        counts[mod + "." + name] = counts.get(mod + "." + name, 0) + 1
            except Exception:
                log_failure(f"FAIL {mod}.{name} hypothesis seed={base}\n{traceback.format_exc()}")
        rnd += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--only", action="append")
    ap.add_argument("--asan", action="store_true", help="run the device code built with AddressSanitizer + UBSan (re-executes itself under LD_PRELOAD)")
    a = ap.parse_args()
    if a.asan and not os.environ.get("GOFR_EMU_LIB"):
        import subprocess
        so = "/tmp/libgofr_emu_asan.so"
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                               "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", so,
                               os.path.join(ROOT, "tests", "emu", "emu_serve.cpp")])
        env = dict(os.environ, GOFR_EMU_LIB=so, ASAN_OPTIONS="detect_leaks=0",
                   LD_PRELOAD=subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip())
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    deadline = time.time() + a.minutes * 60
    with mp.Manager() as mgr:
        counts = [mgr.dict() for _ in range(a.jobs)]
        ps = [mp.Process(target=worker, args=(j, a.jobs, deadline, a.only, counts[j])) for j in range(a.jobs)]
        for p in ps:
            p.start()
        for p in ps:
            p.join()
        total = {}
        for c in counts:
            for k, v in c.items():
                total[k] = total.get(k, 0) + v
        for k in sorted(total):
            print(f"{total[k]:>9}  {k}")
        print("exit codes", [p.exitcode for p in ps])


if __name__ == "__main__":
    main()
