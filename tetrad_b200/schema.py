"""Project file: same fields, file names and JSON layout as the reference's pydantic ``Project``
(tetrad/src/schema.py:32-114) so that a ``{name}.json`` written by either side loads in the other."""
from __future__ import annotations

from pathlib import Path

import numpy as np
from pydantic import BaseModel, Field, computed_field, field_validator


class RNGStateModel(BaseModel):
    """Serialised numpy bit-generator state (reference schema.py:13-29)."""
    bit_generator: str
    state: dict

    @staticmethod
    def from_rng(rng: np.random.Generator) -> "RNGStateModel":
        return RNGStateModel(bit_generator=rng.bit_generator.__class__.__name__,
                             state=rng.bit_generator.state)

    def to_rng(self) -> np.random.Generator:
        rng = np.random.default_rng()
        rng.bit_generator.state = self.state
        return rng


class Project(BaseModel):
    version: str
    # inputs by user (schema.py:37-43)
    data: Path
    name: str
    workdir: Path
    subsample_snps: bool
    weights: int = 0
    nquartets: int = 0
    random_seed: int | None = 0
    # not a field of the reference (its loader ignores unknown keys): how random quartet samples are drawn,
    # "device" (keyed permutation on the GPU, scales with the number of GPUs) or "numpy" (the reference's host draw)
    quartet_sampler: str = "device"
    # checkpointing (schema.py:46-47)
    bootstrap_idx: int = 0
    bootstrap_rng: RNGStateModel | None = None
    # filled by write_database (schema.py:52-56)
    nsnps: int = 0
    nsamples: int = 0
    nqrts: int = 0
    nqrts_total: int = 0
    samples: dict[int, str] = Field(default_factory=dict, repr=False)

    @field_validator("workdir", mode="after")
    @classmethod
    def validate_workdir(cls, value: Path | str) -> Path:
        value = Path(value).expanduser().resolve()
        value.mkdir(exist_ok=True, parents=True)
        return value

    @computed_field
    def json_file(self) -> Path:
        return self.workdir / f"{self.name}.json"

    @computed_field
    def database_file(self) -> Path:
        return self.workdir / f"{self.name}.database.hdf5"

    @computed_field
    def qrts_file(self) -> Path:
        return self.workdir / f"{self.name}.quartets_{self.bootstrap_idx}.tsv"

    @computed_field
    def boots_file(self) -> Path:
        return self.workdir / f"{self.name}.boots.nwk"

    @computed_field
    def best_file(self) -> Path:
        return self.workdir / f"{self.name}.best_tree.nwk"

    @computed_field
    def consensus_file(self) -> Path:
        return self.workdir / f"{self.name}.majority_rule_consensus.nwk"

    @computed_field
    def qmc_in_file(self) -> Path:
        return self.workdir / f"{self.name}.qmc_in.txt"

    @computed_field
    def qmc_out_file(self) -> Path:
        return self.workdir / f"{self.name}.qmc_out.txt"

    def __str__(self):
        return self.model_dump_json(indent=2)

    def save_json(self) -> None:
        with open(self.json_file, "w") as out:
            out.write(self.model_dump_json(indent=2))

    @classmethod
    def load_json(cls, json_file: Path) -> "Project":
        with open(json_file, "r") as indata:
            return cls.model_validate_json(indata.read())
