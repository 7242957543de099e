// scar_comp.h — Sequence_Magic.rcomp's translate table (reference Valkyries/Sequence_Magic.py:97-118): the complement of
// one byte; unmapped bytes pass.  An involution.  Host code; device code uses a 256-byte table filled from it.
#pragma once
#include <stdint.h>

inline uint8_t scar_comp(uint8_t c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'M': return 'K'; case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M';
    case 'V': return 'B'; case 'H': return 'D'; case 'D': return 'H'; case 'B': return 'V';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'm': return 'k'; case 'r': return 'y'; case 'y': return 'r'; case 'k': return 'm';
    case 'v': return 'b'; case 'h': return 'd'; case 'd': return 'h'; case 'b': return 'v';
    default: return c;
    }
}
