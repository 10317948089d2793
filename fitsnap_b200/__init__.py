"""fitsnap_b200 -- B200-native (sm_100a) implementation of FitSNAP's linear-fit hot path.

Only the path is here: descriptor-row assembly into the design matrix A and the weighted
least-squares / ridge solve, behind FitSNAP's own Solver / Calculator plugin API.
See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"
