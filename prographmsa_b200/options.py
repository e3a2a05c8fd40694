"""Explicit mirror of the reference's global ``cmdlineopts`` (main.h:37-82).

Every field that the two hot paths read deep inside their loops
(GraphAlign.h:103-130, DistanceFactoryML.h:98-99,148,177, ModelFactory.h:57-126,
CleanedGraph.h:46, GraphAlign.h:677-680) is an explicit value here and an explicit
struct field in the C ABI (include/pgm_b200.h).  Defaults are main.cpp:63-100,141,159.
"""
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class Options:
    indel_rate: float = 0.0093359375        # -g
    gapext_prob: float = 0.6119140625       # -e
    end_indel_prob: float = 0.12            # -E
    edge_halflife: float = 0.3              # -l
    altsplice_prob: float = 0.328125        # -s
    cutoff_dist: float = 2.2                # -x
    min_dist: float = 0.05                  # -d
    max_dist: float = 2.2                   # -D
    min_pdist: float = 0.05                 # -p
    max_pdist: float = 0.8                  # -P
    repeat_rate: float = 0.1                # --repeat_indel_rate
    repeatext_prob: float = 0.3             # --repeat_indel_ext
    pseudo_count: float = 1000.0            # -C
    mldist_flag: bool = False               # -m
    mldist_gap_flag: bool = False           # -M
    nwdist_flag: bool = False               # -a
    darwin_flag: bool = False               # -w
    aafreqs_flag: bool = False              # -F

    def for_codon(self) -> "Options":
        """--codon rescales the defaults (main.cpp:225-241)."""
        return replace(self, indel_rate=self.indel_rate / 2.6, edge_halflife=self.edge_halflife * 2.6,
                       max_dist=5.0, cutoff_dist=5.0)
