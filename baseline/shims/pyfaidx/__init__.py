"""Stand-in for pyfaidx: Fasta(path)[i] is the i-th record's sequence."""
from pysam import read_fasta


class Fasta:
    def __init__(self, path):
        self._recs = read_fasta(path)

    def __getitem__(self, i):
        return self._recs[i][1]
