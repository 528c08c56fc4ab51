"""Case builders of the package + their CPU checker (oracle/, test infrastructure): `case.make_oracle()`.

The case classes live in fe_project_b200/cases.py (bench.py and smoke() use them too); this module attaches the oracle
builders so that a test reads `o = case.make_oracle(); d = case.make_driver(o)`."""
from fe_project_b200.cases import (C0, MF, SLIP6, DensityCurrentCase, GlobalPanelCase, GlobalSphereCase,  # noqa: F401
                                   SoundWaveCase, rel_l2)
import oracle_cases

DensityCurrentCase.make_oracle = oracle_cases.make_oracle_regional      # SoundWaveCase inherits it
GlobalPanelCase.make_oracle = oracle_cases.make_oracle_panel
GlobalSphereCase.make_oracle = oracle_cases.make_oracle_sphere
