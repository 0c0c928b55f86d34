"""``Tetrad(name, data, nquartets, nboots).run()`` -- the object API the README of the reference
advertises (tetrad/deprecated/tetrad.py:98-110 ctor, :526 ``run``; that module no longer imports
in the reference, so it is restated here over the live ``Project`` + ``run_inference`` engine)."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Sequence

from . import __version__
from .run_inference import run_inference
from .schema import Project
from .write_database import write_database


class Tetrad:
    """Quartet species-tree inference on one or more B200 GPUs.

    Parameters follow the reference class: ``name`` (output prefix), ``data`` (ipyrad
    ``.snps.hdf5`` or ``.npz`` with ``snps``/``snpsmap``/``names``), ``workdir``, ``nquartets``
    (0 = all), ``nboots``, ``save_invariants`` (keep the 16x16 count matrices), ``imap`` (not
    implemented by the live reference either; rejected), ``seed``, ``load`` (resume from
    ``{workdir}/{name}.json``).  Extra keywords: ``subsample_snps`` (default True, one SNP per
    locus as in the reference class), ``weights`` (wQMC weighting 0..3, default 1).
    """

    def __init__(self, name: str, data: str, workdir: str = "analysis-tetrad", nquartets: int = 0,
                 nboots: int = 0, save_invariants: bool = False, imap: Optional[Dict[str, Sequence[str]]] = None,
                 seed: Optional[int] = None, load: bool = False, **kwargs):
        if imap is not None:
            raise NotImplementedError("imap sub-sampling is not implemented (nor in the live reference, "
                                      "tetrad/src/cli_init.py:57)")
        self.name, self.nboots, self.save_invariants = name, int(nboots), bool(save_invariants)
        self.kwargs = kwargs
        json_file = Path(workdir).expanduser().resolve() / f"{name}.json"
        if load:
            self.project = Project.load_json(json_file)
        else:
            self.project = Project(version=__version__, data=Path(data), name=name, workdir=Path(workdir),
                                   subsample_snps=bool(kwargs.get("subsample_snps", True)),
                                   weights=int(kwargs.get("weights", 1)), nquartets=int(nquartets),
                                   random_seed=seed)
            write_database(self.project)
            self.project.save_json()
        self.samples = list(self.project.samples.values())

    @property
    def trees(self):
        """Paths of the result files (reference names, tetrad/src/schema.py:65-95)."""
        p = self.project
        return dict(tree=p.best_file, boots=p.boots_file, cons=p.consensus_file)

    def run(self, force: bool = False, quiet: bool = False, ipyclient=None, auto: bool = False,
            show_cluster: bool = False, device: int = 0):
        """Infer the full tree and ``nboots`` bootstrap replicates.  ``ipyclient``, ``auto`` and
        ``show_cluster`` are accepted for signature compatibility (deprecated/tetrad.py:526):
        there is no ipyparallel cluster, the replicates run on the GPU(s) of this process group."""
        if force:
            self.project.bootstrap_idx = 0
            self.project.bootstrap_rng = None
            if self.project.boots_file.exists():
                self.project.boots_file.unlink()
        run_inference(self.project, ncores=0, nboots=self.nboots, device=device, save_invariants=self.save_invariants)
        return self
