"""`python -m plenoctree_b200.octree.task_manager tasks.json --gpus "0 1 2 3"` — the reference's scene-level driver
(octree/task_manager.py:28-195; SURVEY §8e "scene-level: replicas only"): every task converts one scene
(extraction -> optimization -> evaluation, each its own process pinned to one GPU through CUDA_VISIBLE_DEVICES);
one worker per listed GPU pulls tasks from a shared queue until it is empty.  No collective is involved.

Task file (same keys as the reference's octree/config/*.json):
    {"data_root": ..., "train_root": ...,
     "scenes": [...], "scene_tasks": [{... "{%}" stands for the scene name ...}],
     "tasks": [{"octree_name", "train_dir", "data_dir", "config", "extr_flags", "opt_flags", "eval_flags"}]}
Outputs per task, under <train_dir>/octrees/<octree_name>/: tree.npz (overwritten by the optimised tree unless
--keep_raw, which writes tree_opt.npz beside it) and results.txt:
    <capacity>
    <raw PSNR> <raw SSIM> <raw LPIPS>
    <optimised PSNR> <SSIM> <LPIPS>          (the raw line again when optimisation left no tree)
LPIPS is nan unless the evaluation processes find the LPIPS weight files (plenoctree_b200/nerf/lpips.py).

Differences from the reference, on purpose: commands are argv lists (no shell), the metrics are found by pattern
("capacity:<used>/<reserved>" in the tree's repr, "Average PSNR <p> SSIM <s>") instead of by counting output lines
from the end, a failing task is reported and the worker moves on, and --dry_run prints the commands only.
"""
import argparse
import json
import os
import queue
import re
import subprocess
import sys
import threading

SCENE_MARK = "{%}"


def expand_tasks(spec):
    """all tasks of a task file: the explicit ones plus one per (scene_task, scene); train_dir / data_dir joined
    to their roots."""
    tasks = [dict(t) for t in spec.get("tasks", [])]
    for template in spec.get("scene_tasks", []):
        for scene in spec.get("scenes", []):
            t = dict(template)
            for key in ("data_dir", "train_dir", "octree_name"):
                t[key] = template[key].replace(SCENE_MARK, scene)
            tasks.append(t)
    for t in tasks:
        t["train_dir"] = os.path.join(spec["train_root"], t["train_dir"])
        t["data_dir"] = os.path.join(spec["data_root"], t["data_dir"])
    return tasks


def commands_for(task, keep_raw=False, python=None):
    """-> (store_dir, {"extract": argv, "optimize": argv, "evaluate": argv}, raw tree path, final tree path)."""
    py = [python or sys.executable, "-u", "-m"]
    store = os.path.join(task["train_dir"], "octrees", task["octree_name"])
    raw = os.path.join(store, "tree.npz")
    final = os.path.join(store, "tree_opt.npz") if keep_raw else raw
    common = ["--config", str(task["config"]), "--data_dir", task["data_dir"]]
    cmds = {
        "extract": py + ["octree.extraction", "--train_dir", task["train_dir"], "--is_jaxnerf_ckpt", "--output", raw]
                   + common + list(task.get("extr_flags", [])),
        "optimize": py + ["octree.optimization", "--input", raw, "--output", final] + common
                    + list(task.get("opt_flags", [])),
        "evaluate": py + ["octree.evaluation", "--input", final] + common + list(task.get("eval_flags", [])),
    }
    return store, cmds, raw, final


_CAPACITY = re.compile(r"capacity:(\d+)/\d+")
_METRICS = re.compile(r"Average PSNR\s+([-+0-9.eE]+|nan|inf)\s+SSIM\s+([-+0-9.eE]+|nan)(?:\s+LPIPS\s+([-+0-9.eE]+|nan))?")


def parse_capacity(text):
    m = _CAPACITY.findall(text)
    return int(m[-1]) if m else None


def parse_metrics(text):
    """last "Average PSNR .. SSIM .. [LPIPS ..]" line -> (psnr, ssim, lpips); lpips nan when absent."""
    m = _METRICS.findall(text)
    if not m:
        return None
    p, s, l = m[-1]
    return float(p), float(s), float(l) if l else float("nan")


def convert_one(task, env, keep_raw=False, dry_run=False, run=subprocess.run):
    """one scene: extraction (with its built-in evaluation of the raw tree), optimisation, evaluation; writes
    results.txt.  Returns {"capacity", "raw", "opt"} (None entries where a stage produced nothing)."""
    store, cmds, raw, final = commands_for(task, keep_raw)
    if dry_run:
        for name in ("extract", "optimize", "evaluate"):
            print(" ".join(cmds[name]))
        return {"capacity": None, "raw": None, "opt": None}
    os.makedirs(store, exist_ok=True)
    print("! Extract", task["train_dir"], task["octree_name"], flush=True)
    out = run(cmds["extract"], env=env, check=True, stdout=subprocess.PIPE, text=True).stdout
    capacity, raw_m = parse_capacity(out), parse_metrics(out)
    print("! Optimize", task["train_dir"], task["octree_name"], flush=True)
    run(cmds["optimize"], env=env, check=False)
    opt_m = None
    if os.path.exists(final):
        print("! Eval", task["train_dir"], task["octree_name"], flush=True)
        opt_m = parse_metrics(run(cmds["evaluate"], env=env, check=True, stdout=subprocess.PIPE, text=True).stdout)
    fmt = lambda m: "nan nan nan" if m is None else "%.10f %.10f %.10f" % m
    with open(os.path.join(store, "results.txt"), "w") as f:
        f.write(f"{capacity if capacity is not None else -1}\n{fmt(raw_m)}\n{fmt(opt_m if opt_m is not None else raw_m)}\n")
    print(":", task["octree_name"] or task["train_dir"], "capacity", capacity, "RAW", raw_m, "OPT", opt_m, flush=True)
    return {"capacity": capacity, "raw": raw_m, "opt": opt_m}


def run_all(tasks, gpus, keep_raw=False, dry_run=False, run=subprocess.run):
    """one worker thread per GPU id; every worker's children see only that GPU.  Returns {task index: result | error}."""
    todo = queue.Queue()
    for i, t in enumerate(tasks):
        todo.put((i, t))
    results = {}

    def worker(gpu):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu))
        while True:
            try:
                i, t = todo.get_nowait()
            except queue.Empty:
                return
            try:
                results[i] = convert_one(t, env, keep_raw, dry_run, run)
            except Exception as e:   # noqa: BLE001  one broken scene must not stop the others
                results[i] = {"error": f"{type(e).__name__}: {e}"}
                print("! FAILED", t["train_dir"], results[i]["error"], flush=True)

    threads = [threading.Thread(target=worker, args=(g,), daemon=True) for g in gpus]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    return results


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("task_json", type=str)
    ap.add_argument("--gpus", type=str, required=True, help="space delimited GPU id list (pre CUDA_VISIBLE_DEVICES)")
    ap.add_argument("--keep_raw", action="store_true", help="do not overwrite raw octree (takes extra disk space)")
    ap.add_argument("--dry_run", action="store_true", help="print the commands of every task and exit")
    args = ap.parse_args(argv)
    if os.path.exists(args.task_json):
        with open(args.task_json) as f:
            spec = json.load(f)
    else:
        from ..presets import octree_tasks_preset        # octree/config/{syn_sh16,tt_sh25}.json by base name
        spec = octree_tasks_preset(args.task_json)
        if spec is None:
            raise FileNotFoundError(args.task_json)
    tasks = expand_tasks(spec)
    print(len(tasks), "total tasks")
    if not args.dry_run:
        for t in tasks:
            for key in ("train_dir", "data_dir"):
                if not os.path.exists(t[key]):
                    raise FileNotFoundError(t[key])
    gpus = [int(g) for g in args.gpus.split()]
    print("GPUS:", gpus)
    results = run_all(tasks, gpus, args.keep_raw, args.dry_run)
    failed = [i for i, r in results.items() if "error" in r]
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
