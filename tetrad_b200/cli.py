"""Command line: ``tetrad init DATA -n NAME -w DIR [-q N] [-r SEED] [-x W] [-s]`` and
``tetrad run JSON [-c CORES] [-b BOOTS]`` with the reference's flags (tetrad/src/cli_init.py:52-62,
tetrad/src/cli_run.py:49-53), plus the one-shot form of its README,
``tetrad -i DATA [-n NAME] [-o OUTDIR] [-q N] [-b BOOTS] [-c CORES]`` (README.rst:34-41).
Run as ``python -m tetrad_b200 ...``."""
from __future__ import annotations

import logging
import sys
from argparse import ArgumentParser
from pathlib import Path

from . import __version__
from .schema import Project


def _parser() -> ArgumentParser:
    parser = ArgumentParser("tetrad", description="quartet species tree analysis (B200 hot path)")
    parser.add_argument("-v", "--version", action="version", version=__version__)
    parser.add_argument("--log-level", default="INFO", choices=["DEBUG", "INFO", "WARNING", "EXCEPTION"])
    sub = parser.add_subparsers(dest="subcommand")
    init = sub.add_parser("init", help="create a project JSON + database")
    init.add_argument("data", type=Path, help="A SNPS.HDF5 (or .npz) file")
    init.add_argument("-n", "--name", type=str, required=True)
    init.add_argument("-w", "--workdir", type=Path, default=".")
    init.add_argument("-q", "--nquartets", type=float, default=0)
    init.add_argument("-r", "--random-seed", type=int)
    init.add_argument("-x", "--weights", type=int, default=1)
    init.add_argument("-s", "--subsample-snps", action="store_true")
    init.add_argument("--quartet-sampler", choices=["device", "numpy"], default="device",
                      help="random quartet samples: keyed permutation on the GPU (default; scales with the number of GPUs) "
                           "or the reference's host draw rng.choice (same sample as the reference for the same seed)")
    run = sub.add_parser("run", help="run tree inference on a project")
    run.add_argument("json", type=Path)
    run.add_argument("-c", "--cores", type=int, default=0, help="ignored: work runs on the GPU(s)")
    run.add_argument("-b", "--boots", type=int, default=0)
    run.add_argument("--device", type=int, default=0)
    return parser


def _legacy(argv) -> int:
    ap = ArgumentParser("tetrad")
    ap.add_argument("-i", dest="data", type=Path, required=True)
    ap.add_argument("-n", dest="name", default="test")
    ap.add_argument("-o", dest="outdir", type=Path, default=Path("analysis-tetrad"))
    ap.add_argument("-q", dest="nquartets", type=float, default=0)
    ap.add_argument("-b", dest="boots", type=int, default=0)
    ap.add_argument("-c", dest="cores", type=int, default=0)
    ap.add_argument("-r", dest="seed", type=int)
    ap.add_argument("-l", dest="linkage", type=int, help="accepted and ignored (README.rst:41; not implemented by the reference)")
    args = ap.parse_args(argv)
    from .tetrad import Tetrad
    Tetrad(args.name, str(args.data), workdir=str(args.outdir), nquartets=int(args.nquartets), nboots=args.boots,
           seed=args.seed).run()
    return 0


def _init_distributed(device: int):
    """Under ``torchrun`` (one process per GPU) every rank loads the same project, takes the same draws from the
    project's generator, resolves its contiguous share of every replicate's quartets and rank 0 gathers them over
    NCCL and writes the files (``sharding.resolve_sharded``; the reference farms chunks out over ipyparallel,
    run_inference.py:184-251).  Returns (device index, whether a process group was created here)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return device, False
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return int(os.environ.get("LOCAL_RANK", device)), False
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if dist.get_rank() != 0:
        logging.getLogger().setLevel(logging.WARNING)       # one log stream, rank 0's
    return local, True


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "-i" or (len(argv) > 1 and "-i" in argv and argv[0] not in ("init", "run")):
        return _legacy(argv)
    args = _parser().parse_args(argv)
    logging.basicConfig(level=getattr(logging, args.log_level if args.log_level != "EXCEPTION" else "ERROR"),
                        format="%(asctime)s %(levelname)s | %(name)s | %(message)s", stream=sys.stderr)
    if args.subcommand == "init":
        from .write_database import write_database
        proj = Project(version=__version__, data=args.data, name=args.name, workdir=args.workdir,
                       subsample_snps=args.subsample_snps, weights=args.weights, nquartets=int(args.nquartets),
                       random_seed=args.random_seed, quartet_sampler=args.quartet_sampler)
        write_database(proj)
        proj.save_json()
        return 0
    if args.subcommand == "run":
        from .run_inference import run_inference
        device, group = _init_distributed(args.device)
        try:
            run_inference(Project.load_json(args.json), args.cores, args.boots, device=device)
        finally:
            if group:
                import torch.distributed as dist
                dist.destroy_process_group()
        return 0
    _parser().print_help()
    return 2


if __name__ == "__main__":
    raise SystemExit(main())
