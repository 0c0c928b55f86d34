"""``tetrad init``: build the working database from an input SNP file.  Mirrors the reference's
tetrad/src/write_database.py (``get_nquartets`` :81-107, ``init_database`` :110-176,
``write_database`` :179-195)."""
from __future__ import annotations

import logging
from math import comb
from pathlib import Path

import numpy as np

from . import database, synth

logger = logging.getLogger("tetrad_b200")


def get_nquartets(nsamples: int, nquartets: int) -> tuple[int, int]:
    """(quartets to sample, total quartets); reference write_database.py:81-107."""
    rough = int(nsamples ** 2.8)
    total = int(comb(nsamples, 4))
    assert total < 4_294_967_295, "max possible quartets exceeded."
    if not nquartets:
        logger.info(f"quartet sampler [full]: {total}/{total}")
        return total, total
    nquartets = int(nquartets)
    if nquartets < rough:
        logger.warning(f"nquartets is low ({nquartets}/{total}), consider raising to {rough} or higher")
        return min(nquartets, total), total
    if nquartets > total:
        logger.info(f"quartet sampler [full]: {total}/{total}")
        return total, total
    logger.info(f"quartet sampler [random]: {nquartets}/{total}")
    return nquartets, total


def init_database(data: Path, out: Path, rng) -> dict:
    """Write ``spans``, ``tmpmap``, ``seqarr``, ``tmparr`` to ``out`` (reference :110-176) and
    return (names, nsnps)."""
    snps, snpsmap, names = database.read_snps_input(data)
    seed = int(np.random.default_rng(rng).integers(2**31))
    arrays = synth.prepare_database(np.asarray(snps), np.asarray(snpsmap), seed=seed)
    nloci = arrays["spans"].shape[0]
    logger.info(f"max unlinked SNPs per quartet [nloci]: {nloci}")
    database.write_arrays(out, arrays, as_npz=database._kind(Path(data)) == "npz")
    logger.info(f"wrote database file to {out}")
    return dict(names=names, nsnps=int(snps.shape[1]))


def write_database(project) -> None:
    """Fill the project's derived fields and write its database (reference :179-195).  A project created without
    ``-r`` gets a concrete seed drawn from the OS here and kept in its JSON: ``run`` then reproduces, and the ranks of
    a multi-GPU run (one process each, every one drawing from its own copy of the project's generator) agree."""
    if project.random_seed is None:
        project.random_seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0] >> np.uint64(1))
        logger.info(f"no random seed given: drew {project.random_seed} (recorded in the project file)")
    info = init_database(project.data, project.database_file, project.random_seed)
    nsamples = len(info["names"])
    nqrts, nqrts_total = get_nquartets(nsamples, project.nquartets)
    project.nqrts = nqrts
    project.nqrts_total = nqrts_total
    project.nsamples = nsamples
    project.nsnps = info["nsnps"]
    project.samples = dict(enumerate(info["names"]))
