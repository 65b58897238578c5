// Host-only checks of the layout arithmetic shared by the packer, the search kernels and the parity hooks
// (hpg_variant_b200/csrc/epi_device.cuh, epi_kernels.cuh).  Built with nvcc as a plain host program and run by
// tests/test_layout_native.py on the CPU: no kernel is launched.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
#include "../../hpg_variant_b200/csrc/epi_kernels.cuh"

using namespace hpgv;

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { fails++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

int main() {
    // ---- tri layout: every logical bit (block, plane, word, bit) has its own physical bit, inside the row, outside the marginals
    for (int nblocks = 4; nblocks <= 24; nblocks += 4) {
        const int rw = tri_row_words(nblocks), ngroups = nblocks / 4, ntail = (nblocks + 7) / 8;
        CHECK(rw % 4 == 0 && (rw / 4) % 2 == 1, "tri_row_words(%d) = %d is not an odd number of 16-byte groups", nblocks, rw);
        CHECK(tri_marg_off(nblocks, 0) == ngroups * 36 + ntail * 4, "marginals do not follow the tails");
        CHECK(tri_marg_off(nblocks, ngroups - 1) + 4 <= rw, "marginals of %d blocks leave the row", nblocks);
        std::set<long> used;
        for (int b = 0; b < nblocks; b++)
            for (int g = 0; g < 3; g++) {
                for (int w = 0; w < 3; w++)
                    for (int bit = 0; bit < 32; bit++) {
                        const long phys = (long) tri_word_off(b, g, w) * 32 + bit;
                        CHECK(tri_word_off(b, g, w) < ngroups * 36, "main word outside the groups");
                        CHECK(used.insert(phys).second, "tri: two logical bits share physical bit %ld", phys);
                    }
                for (int bit = 0; bit < 4; bit++) {
                    const int off = tri_tail_off(nblocks, b, g);
                    CHECK(off >= ngroups * 36 && off < tri_marg_off(nblocks, 0), "tail word outside the tail area");
                    const long phys = (long) off * 32 + tri_tail_shift(b) + bit;
                    CHECK(used.insert(phys).second, "tri: tail bit collides at %ld", phys);
                }
                // the tail's nibble is the one whose count lands in the byte counter of its block:
                // counter word k = b / 4 keeps block q = b % 4 in byte group_shift(q) / 8; tail word m = k / 2 splits into
                // even nibbles -> counter word 2m, odd nibbles -> counter word 2m + 1 (tail_count_acc)
                const int k = b / 4, q = b % 4, nib = tri_tail_shift(b) / 4;
                CHECK(tri_tail_shift(b) % 4 == 0 && nib / 2 == (int) group_shift(q) / 8 && (nib & 1) == (k & 1),
                      "tail nibble of block %d does not match its byte counter", b);
                CHECK(tri_tail_off(nblocks, b, g) - tri_tail_off(nblocks, b, 0) == g, "tail planes are not adjacent");
            }
    }
    // group_shift: the four blocks of a group own the four bytes of a counter word, (A_2k, A_2k+1, U_2k, U_2k+1)
    {
        std::set<unsigned> bytes;
        for (int q = 0; q < 4; q++) bytes.insert(group_shift(q));
        CHECK(bytes.size() == 4 && group_shift(0) == 0 && group_shift(2) == 8 && group_shift(1) == 16 && group_shift(3) == 24, "group_shift");
    }
    // ---- linear combination indices: pair_index / triple_index enumerate lexicographic order without gaps
    for (int n : {2, 3, 7, 40}) {
        uint64_t expect = 0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                CHECK(pair_index(n, i, j) == expect, "pair_index(%d, %d, %d)", n, i, j);
                expect++;
            }
        CHECK(expect == choose2(n), "choose2(%d)", n);
        expect = 0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++)
                for (int k = j + 1; k < n; k++) {
                    CHECK(triple_index(n, i, j, k) == expect, "triple_index(%d, %d, %d, %d)", n, i, j, k);
                    expect++;
                }
        CHECK(expect == choose3(n), "choose3(%d)", n);
    }
    CHECK(choose3(3000000) == 3000000ull * 2999999ull / 2 * 2999998ull / 3, "choose3 overflows early");
    // ---- shared-memory maps: regions are ordered, aligned and sized by the same function on host and device
    {
        FoldLayout fl{};
        fl.F = 10; fl.single = 1; fl.tri = 1; fl.bw = 4; fl.nblocks = 20; fl.cb = 20; fl.nchunks = 1; fl.row_words = tri_row_words(20);
        for (int ns = 2; ns <= 3; ns++) {
            const SmemMap m = search_smem_map(fl, kTriWarps + kTileJ, 9, kTriWarps * 32, 50, true, ns);
            CHECK(m.stage0 % 128 == 0 && m.stage_bytes % 128 == 0, "stages are not 128-byte aligned");
            CHECK(m.counters == m.stage0 + ns * m.stage_bytes && m.desc > m.counters && m.lists >= m.desc && m.lists % 16 == 0, "region order");
            CHECK(m.total == m.lists + (size_t) 10 * 50 * sizeof(Cand), "lists size");
        }
        CHECK(search_smem_map(fl, kTriWarps + kTileJ, 9, kTriWarps * 32, 50, true, 2).total <= 227 * 1024, "the c2 shape must fit one SM");
        const PackSmem p = pack_smem_map(20 * 128, 2000, fl);
        CHECK(p.ofs == 20 * 128 * 4 && p.row % 16 == 0 && p.out % 16 == 0 && p.total == p.out + (size_t) kPackRows * fl.row_words * 4, "pack map");
        CHECK(hist_coarse_bins(1001) == 32 && hist_coarse_bins(32) == 1 && hist_coarse_bins(33) == 2, "hist_coarse_bins");
        CHECK(counter_stride(9, 5) % 2 == 1 && counter_stride(27, 5) % 2 == 1, "counter stride must be odd");
        CHECK(merge_smem_bytes(50) >= (size_t) kMergeSort * 20 && merge_smem_bytes(4096) >= (size_t) 4096 * 20 + 16 && merge_smem_bytes(4096) <= 227 * 1024
              && kMergeSort >= 148 * 50 && (kMergeSort & (kMergeSort - 1)) == 0, "merge smem");
    }
    printf(fails ? "%d check(s) failed\n" : "layout checks ok\n", fails);
    return fails ? 1 : 0;
}
