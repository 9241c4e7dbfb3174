"""Multi-host launcher: role of ``python3 -m torch_xla.distributed.xla_dist --tpu=... --env ... -- python3 run_vit_training.py``
in the reference (README.md:99-118): fan one command out to every host over SSH, forward environment variables, and
give every host its node rank.  On GPUs the per-host process group is started by ``torchrun``; this tool only builds
and dispatches the per-host command lines.

    python -m vit_10b_fsdp_example_b200.pod_launch --hosts gpu-a,gpu-b --nproc-per-node 8 \
        --env NCCL_DEBUG=WARN --env PYTHONUNBUFFERED=1 -- run_vit_training.py --data_dir /datasets/imagenet-1k ...

* rank 0's host (first in ``--hosts``) is the rendezvous endpoint (``--master-port``, default 29500);
* ``--restart`` (the analogue of ``--restart-tpuvm-pod-server``) first stops the python processes recorded in this
  tool's pid file on every host (exact PIDs, never a pattern), so a crashed job does not block the ports;
* across hosts the engine selects the NCCL backend automatically (the symmetric-memory kernels are single-box);
* a non-zero exit on any host terminates the others (same contract as the reference: restart with ``--resume_epoch``).
"""
from __future__ import annotations

import argparse
import shlex
import subprocess
import sys
from typing import Dict, List

PID_FILE = "/tmp/vit_fsdp_pod_launch.pid"


def build_host_command(node_rank: int, hosts: List[str], nproc: int, port: int, env: Dict[str, str], script: List[str],
                       workdir: str, python: str) -> str:
    """The shell command one host runs (a torchrun node joining the job)."""
    exports = " ".join(f"{k}={shlex.quote(v)}" for k, v in sorted(env.items()))
    torchrun = [python, "-m", "torch.distributed.run", f"--nnodes={len(hosts)}", f"--node-rank={node_rank}",
                f"--nproc-per-node={nproc}", f"--master-addr={hosts[0]}", f"--master-port={port}", *script]
    cmd = " ".join(shlex.quote(c) for c in torchrun)
    return f"cd {shlex.quote(workdir)} && echo $$ > {PID_FILE} && exec env {exports} {cmd}".replace("env  ", "env ")


def build_restart_command() -> str:
    """Stop the job this tool started earlier on that host: exact PID (the torchrun agent, which owns its workers)."""
    return f"if [ -f {PID_FILE} ]; then kill $(cat {PID_FILE}) 2>/dev/null; rm -f {PID_FILE}; fi; true"


def parse(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--hosts", required=True, help="comma-separated host names; the first one hosts the rendezvous")
    ap.add_argument("--nproc-per-node", type=int, default=8)
    ap.add_argument("--master-port", type=int, default=29500)
    ap.add_argument("--env", action="append", default=[], metavar="K=V", help="forwarded to every process (repeatable)")
    ap.add_argument("--workdir", default=".", help="directory of the checkout on every host")
    ap.add_argument("--python", default="python3")
    ap.add_argument("--ssh", default="ssh -o BatchMode=yes -o StrictHostKeyChecking=accept-new")
    ap.add_argument("--restart", action="store_true", help="stop a previous launch of this tool on every host first")
    ap.add_argument("--dry-run", action="store_true", help="print the per-host commands instead of running them")
    ap.add_argument("script", nargs=argparse.REMAINDER, help="-- script.py and its flags")
    args = ap.parse_args(argv)
    if args.script and args.script[0] == "--":
        args.script = args.script[1:]
    if not args.script:
        ap.error("nothing to launch: give the training script after --")
    return args


def main(argv=None) -> int:
    args = parse(argv)
    hosts = [h for h in args.hosts.split(",") if h]
    env = dict(kv.split("=", 1) for kv in args.env)
    ssh = shlex.split(args.ssh)
    plans = [(h, build_host_command(i, hosts, args.nproc_per_node, args.master_port, env, args.script, args.workdir,
                                    args.python)) for i, h in enumerate(hosts)]
    if args.dry_run:
        for h, c in plans:
            if args.restart:
                print(f"[{h}] {build_restart_command()}")
            print(f"[{h}] {c}")
        return 0
    if args.restart:
        for h in hosts:
            subprocess.run([*ssh, h, build_restart_command()], check=False)
    procs = [(h, subprocess.Popen([*ssh, h, c])) for h, c in plans]
    rc = 0
    try:
        while procs and rc == 0:
            for h, p in list(procs):
                r = p.poll()
                if r is None:
                    continue
                procs.remove((h, p))
                if r != 0:
                    print(f"[pod_launch] host {h} exited with {r}: stopping the other hosts", file=sys.stderr)
                    rc = r
            if procs and rc == 0:
                try:
                    procs[0][1].wait(timeout=1.0)
                except subprocess.TimeoutExpired:
                    pass
    finally:
        for _, p in procs:  # exact processes this tool started
            p.terminate()
    return rc


if __name__ == "__main__":
    sys.exit(main())
