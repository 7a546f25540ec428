"""Fixed r-torsion generators of BLS12-381 G1 / G2 used to build public parameters for benchmarks and tests.

The reference samples g and h with `G1Projective::rand(rng)` / `G2Projective::rand(rng)`
(/root/reference/src/commitment/setup.rs:28-31); any generators of the prime-order subgroups give an
equally valid PublicParameter.  G1 is the standard BLS12-381 generator; the G2 point was derived by
hash-and-clear-cofactor from x = (1, 1) (tests/golden/make_golden.py re-derives both).  Values are
canonical integers; the arrays exported below are arkworks' Montgomery limbs (6 x u64 per Fq).
"""
import numpy as np

FQ_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
_R = 1 << 384

G1_X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1_Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
G2_X = (0x04d1cc4ad56b68cdb595adb46cad2cc82e3d0da9a75ef283b6bbd91df14533e1a45128ec26f8ab25072da969d7628b70,
        0x13a471d5149813b306fe76921cff7bb8d5c03fdc24a613f3e7a7fb8deb8097699751485a0bd2ad391718aaa4419ce75b)
G2_Y = (0x0a3d002cac5c50eb9e97e8b62ca30ffc5bf5aaacec121cdb63e19a5e358c4804439edb98366c02fd2840c7b9004f8b99,
        0x1834907430540701fa8aa597f79e63960ec77037a7d9a06606c4c58bd8019969edabb81b77fae18489a80d47bab79d25)


def fq_mont_limbs(v):
    m = v * _R % FQ_MOD
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)]


G1_GENERATOR = np.array(fq_mont_limbs(G1_X) + fq_mont_limbs(G1_Y), dtype=np.uint64)
G2_GENERATOR = np.array(fq_mont_limbs(G2_X[0]) + fq_mont_limbs(G2_X[1]) + fq_mont_limbs(G2_Y[0]) + fq_mont_limbs(G2_Y[1]), dtype=np.uint64)
