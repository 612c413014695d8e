/* TEST INFRASTRUCTURE (oracle) — plain-C restatement of the reference's only native component, the BPR negative
 * sampler: code/sources/sampling.cpp:22-25 (randint_), :27-56 (sample_negative), :58-86 (sample_negative_ByUser),
 * :88-91 (seed).  Same glibc rand() stream, same draw order; the per-user positive lists arrive as a CSR
 * (indptr int64[user_num+1], items int32) instead of vector<vector<int>>.  Pinned against the reference itself by
 * tests/golden/sampler.npz (made by compiling sampling.cpp as it lies, oracle/gen_golden.py::sampler_golden).
 * bench.py --impl reference draws its triples with this file, so the reference arm loads none of the product's code. */
#include <stdint.h>
#include <stdlib.h>

void oracle_sampler_seed(unsigned int seed) { srand(seed); }                 /* sampling.cpp:88-91 */
int oracle_randint(int end) { return rand() % end; }                          /* sampling.cpp:22-25 */

static int contains(const int32_t* pos, int64_t n, int v) {                   /* std::find over the user's positives */
    for (int64_t j = 0; j < n; ++j) if (pos[j] == v) return 1;
    return 0;
}

static void one_row(int user, const int32_t* pos, int64_t npos, int item_num, int neg_num, int32_t* o) {
    o[0] = user;
    o[1] = pos[rand() % npos];
    for (int idx = 2; idx < neg_num + 2; ++idx) {
        int neg;
        do { neg = rand() % item_num; } while (contains(pos, npos, neg));
        o[idx] = neg;
    }
}

/* sampling.cpp:27-56.  out int32[user_num * (train_num / user_num) * (2 + neg_num)]; returns rows written, <0 on a user
 * without positives (the reference divides by zero there). */
int64_t oracle_sample_negative(int user_num, int item_num, int64_t train_num, const int64_t* indptr, const int32_t* items,
                               int neg_num, int32_t* out) {
    const int64_t per_user = train_num / user_num;
    const int row = neg_num + 2;
    for (int user = 0; user < user_num; ++user) {
        const int64_t npos = indptr[user + 1] - indptr[user];
        if (npos <= 0 && per_user > 0) return -2;
        for (int64_t pair = 0; pair < per_user; ++pair)
            one_row(user, items + indptr[user], npos, item_num, neg_num, out + ((int64_t)user * per_user + pair) * row);
    }
    return (int64_t)user_num * per_user;
}

/* sampling.cpp:58-86.  out int32[n_users_listed * (2 + neg_num)]. */
int64_t oracle_sample_negative_by_user(const int32_t* users, int64_t n_listed, int item_num, const int64_t* indptr,
                                       const int32_t* items, int neg_num, int32_t* out) {
    const int row = neg_num + 2;
    for (int64_t i = 0; i < n_listed; ++i) {
        const int user = users[i];
        const int64_t npos = indptr[user + 1] - indptr[user];
        if (npos <= 0) return -2;
        one_row(user, items + indptr[user], npos, item_num, neg_num, out + i * row);
    }
    return n_listed;
}
