"""
alignment_oracle.py — CPU restatement of AlignmentProcessing.alignment_processing
(reference scarmapper/AlignmentProcessing.pyx:18-149).  TEST INFRASTRUCTURE ONLY.

A literal pure-Python statement of the walk (small cases only), pinned by tests/test_alignment.py against the
reference's own module compiled in place (oracle/_ref/AlignmentProcessing.*.so, oracle/build_ref.py) on randomised
gapped alignments, and against the committed golden vectors tests/golden/alignment_cases.json.gz (generated from the
compiled reference by tests/golden/make_alignment_golden.py).
"""


def alignment_processing(cutposition, gap_con_len, gap_genome_len, gapped_genomic, gapped_consensus, summary_data):
    left_deletion = right_deletion = left_insertion = right_insertion = 0
    lower_limit = 0.10 * gap_genome_len                      # :34-35 (doubles)
    upper_limit = 0.90 * gap_genome_len
    indel = False
    microhomology, insertion, snv = [], [], []
    for position, (c, g) in enumerate(zip(gapped_consensus, gapped_genomic)):      # :45
        temp_homology = tmp_snv = tmp_insertion = ""
        if c == g or position < lower_limit or position > upper_limit:             # :50
            indel = False
        elif c == "-":                                                             # :54-86
            if position <= cutposition:
                left_deletion += 1
            else:
                right_deletion += 1
            if not indel:
                new_position = position
                while new_position < gap_con_len and gapped_consensus[new_position] == "-":
                    new_position += 1
                left_step, right_step = position, new_position
                while True:
                    if right_step >= gap_con_len or left_step >= new_position or gapped_consensus[right_step] == "N":
                        break
                    if gapped_genomic[left_step] == gapped_consensus[right_step]:
                        temp_homology += gapped_genomic[left_step]
                    else:
                        break
                    right_step += 1
                    left_step += 1
            indel = True
            if temp_homology:
                microhomology.append(temp_homology)
        elif g == "-" and c != "N":                                                # :89-108
            if position <= cutposition:
                left_insertion += 1
            else:
                right_insertion += 1
            if not indel:
                new_position = position
                while new_position < len(gapped_genomic) and gapped_genomic[new_position] == "-" and \
                        gapped_consensus[new_position] != "N":
                    tmp_insertion += gapped_consensus[new_position]
                    new_position += 1
                if tmp_insertion:
                    insertion.append(tmp_insertion)
            indel = True
        elif c != g and c != "N" and g != "N":                                     # :110-123
            if not indel:
                new_position = position
                while new_position < gap_con_len and gapped_genomic[new_position] != gapped_consensus[new_position] \
                        and gapped_genomic[new_position] != "-" and gapped_consensus[new_position] != "N" \
                        and gapped_consensus[new_position] != "-":
                    tmp_snv += gapped_consensus[new_position]
                    new_position += 1
                if tmp_snv:
                    snv.append(tmp_snv)
            indel = True
    if microhomology:                                                              # :130-147
        summary_data[8] += 1
    results = [left_deletion, right_deletion, left_insertion, right_insertion, microhomology if microhomology else "",
               insertion if insertion else "", snv if snv else "", gapped_consensus, gapped_genomic]
    if left_deletion > 0:
        summary_data[2] += 1
    if right_deletion > 0:
        summary_data[3] += 1
    if left_insertion > 0:
        summary_data[5] += 1
    if right_insertion > 0:
        summary_data[6] += 1
    return results, summary_data


def random_alignment(rng, n=None):
    """A gapped alignment (genomic, consensus) of one length with deletions, insertions, mismatch runs and N's,
    events near both ends (outside the 10-90 % window) included."""
    n = n or rng.randint(20, 260)
    ref = [rng.choice("ACGT") for _ in range(n)]
    g, c = [], []
    i = 0
    while i < n:
        u = rng.random()
        if u < 0.04:                                   # deletion in the consensus, sometimes with microhomology after it
            k = rng.randint(1, 12)
            seg = ref[i:i + k]
            g += seg
            c += ["-"] * len(seg)
            i += len(seg)
            if rng.random() < 0.5 and seg:
                m = rng.randint(1, min(4, len(seg)))
                for q in range(m):
                    g.append(ref[i + q] if i + q < n else "A")
                    c.append(seg[q])
                i += m
        elif u < 0.07:                                 # insertion in the consensus
            k = rng.randint(1, 8)
            g += ["-"] * k
            c += [rng.choice("ACGTN") if rng.random() < 0.1 else rng.choice("ACGT") for _ in range(k)]
        elif u < 0.10:                                 # mismatch run
            k = rng.randint(1, 4)
            for q in range(k):
                if i + q < n:
                    g.append(ref[i + q])
                    c.append(rng.choice([x for x in "ACGT" if x != ref[i + q]]))
            i += k
        elif u < 0.12:
            g.append(ref[i]); c.append("N"); i += 1
        elif u < 0.13:
            g.append("N"); c.append(ref[i]); i += 1
        else:
            g.append(ref[i]); c.append(ref[i]); i += 1
    return "".join(g), "".join(c)
