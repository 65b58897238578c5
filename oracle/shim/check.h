/*
 * check.h -- minimal stand-in for libcheck (not installed in this image), so
 * that the reference's own unit tests (test/test_epistasis_model.c, ...) can
 * be compiled unchanged from /root/reference against (a) the reference's
 * sources and (b) the oracle restatement.  Test infrastructure only.
 *
 * Every START_TEST body becomes a function; failures are counted (not
 * aborted on) and reported per test.
 */
#ifndef ORACLE_CHECK_SHIM_H
#define ORACLE_CHECK_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*ck_test_fn)(void);
typedef struct { const char *name; ck_test_fn fn; } ck_test;
typedef struct TCase { const char *name; ck_test tests[64]; int n; } TCase;
typedef struct Suite { const char *name; TCase *cases[32]; int n; } Suite;
typedef struct SRunner { Suite *suite; int failed; } SRunner;
enum print_output { CK_SILENT, CK_MINIMAL, CK_NORMAL, CK_VERBOSE };

static int ck_current_failures;
static int ck_total_checks;

#define START_TEST(name) static void name(void)
#define END_TEST

static inline void ck_report(const char *file, int line, const char *msg, ...) {
    ck_current_failures++;
    fprintf(stderr, "  FAIL %s:%d: %s\n", file, line, msg ? msg : "");
}
#define fail_if(cond, ...) do { ck_total_checks++; if (cond) ck_report(__FILE__, __LINE__, ##__VA_ARGS__, NULL); } while (0)
#define fail_unless(cond, ...) do { ck_total_checks++; if (!(cond)) ck_report(__FILE__, __LINE__, ##__VA_ARGS__, NULL); } while (0)
#define ck_assert(cond) fail_unless(cond, #cond)
#define ck_assert_int_eq(a, b) fail_unless((a) == (b), #a " == " #b)
#define ck_assert_msg(cond, ...) fail_unless(cond, __VA_ARGS__)

static inline TCase *tcase_create(const char *name) { TCase *t = calloc(1, sizeof(TCase)); t->name = name; return t; }
#define tcase_add_test(tc, f) do { (tc)->tests[(tc)->n].name = #f; (tc)->tests[(tc)->n].fn = f; (tc)->n++; } while (0)
static inline void tcase_add_unchecked_fixture(TCase *t, void (*s)(void), void (*e)(void)) { (void) t; (void) e; if (s) s(); }
static inline void tcase_add_checked_fixture(TCase *t, void (*s)(void), void (*e)(void)) { (void) t; (void) e; if (s) s(); }
static inline void tcase_set_timeout(TCase *t, int s) { (void) t; (void) s; }
static inline Suite *suite_create(const char *name) { Suite *s = calloc(1, sizeof(Suite)); s->name = name; return s; }
static inline void suite_add_tcase(Suite *s, TCase *t) { s->cases[s->n++] = t; }
static inline SRunner *srunner_create(Suite *s) { SRunner *r = calloc(1, sizeof(SRunner)); r->suite = s; return r; }
static inline void srunner_run_all(SRunner *r, int mode) {
    (void) mode;
    int ntests = 0;
    for (int c = 0; c < r->suite->n; c++) {
        TCase *t = r->suite->cases[c];
        for (int i = 0; i < t->n; i++) {
            ck_current_failures = 0;
            t->tests[i].fn();
            ntests++;
            fprintf(stderr, "%s: %s: %s\n", ck_current_failures ? "FAILED" : "ok", t->name, t->tests[i].name);
            if (ck_current_failures) r->failed++;
        }
    }
    fprintf(stderr, "%s: %d tests, %d failed, %d checks\n", r->suite->name, ntests, r->failed, ck_total_checks);
}
static inline int srunner_ntests_failed(SRunner *r) { return r->failed; }
static inline void srunner_free(SRunner *r) { free(r); }

#endif
