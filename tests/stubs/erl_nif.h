/*
 * Minimal stand-in for Erlang/OTP's erl_nif.h: just the declarations elixir/c_src/nxsignal_nif.c
 * uses, with the OTP 27 signatures, so that `gcc -fsyntax-only` can type-check the NIF in an image
 * without a BEAM (tests/test_elixir_boundary.py).  TEST INFRASTRUCTURE: never used to build anything.
 */
#ifndef NXS_STUB_ERL_NIF_H
#define NXS_STUB_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>

typedef struct enif_environment_t ErlNifEnv;
typedef uintptr_t ERL_NIF_TERM;
typedef int64_t ErlNifSInt64;
typedef uint64_t ErlNifUInt64;
typedef struct {
  size_t size;
  unsigned char* data;
  void* ref_bin;
  void* __spare__[2];
} ErlNifBinary;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef struct ErlNifMutex_ ErlNifMutex;
typedef void ErlNifResourceDtor(ErlNifEnv*, void*);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
typedef enum { ERL_NIF_LATIN1 = 1, ERL_NIF_UTF8 = 2 } ErlNifCharEncoding;
#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2
typedef struct {
  const char* name;
  unsigned arity;
  ERL_NIF_TERM (*fptr)(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]);
  unsigned flags;
} ErlNifFunc;

int enif_get_int(ErlNifEnv*, ERL_NIF_TERM, int*);
int enif_get_int64(ErlNifEnv*, ERL_NIF_TERM, ErlNifSInt64*);
int enif_get_double(ErlNifEnv*, ERL_NIF_TERM, double*);
int enif_get_list_length(ErlNifEnv*, ERL_NIF_TERM, unsigned*);
int enif_get_list_cell(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM* head, ERL_NIF_TERM* tail);
int enif_inspect_binary(ErlNifEnv*, ERL_NIF_TERM, ErlNifBinary*);
int enif_get_resource(ErlNifEnv*, ERL_NIF_TERM, ErlNifResourceType*, void** objp);
unsigned char* enif_make_new_binary(ErlNifEnv*, size_t, ERL_NIF_TERM* termp);
ERL_NIF_TERM enif_make_atom(ErlNifEnv*, const char*);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv*);
ERL_NIF_TERM enif_make_int64(ErlNifEnv*, ErlNifSInt64);
ERL_NIF_TERM enif_make_string(ErlNifEnv*, const char*, ErlNifCharEncoding);
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple3(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple5(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_list3(ErlNifEnv*, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_resource(ErlNifEnv*, void* obj);
void* enif_alloc_resource(ErlNifResourceType*, size_t);
void enif_release_resource(void* obj);
ErlNifResourceType* enif_open_resource_type(ErlNifEnv*, const char* module_str, const char* name_str,
                                            ErlNifResourceDtor* dtor, ErlNifResourceFlags flags,
                                            ErlNifResourceFlags* tried);
ErlNifMutex* enif_mutex_create(char* name);
void enif_mutex_destroy(ErlNifMutex*);
void enif_mutex_lock(ErlNifMutex*);
void enif_mutex_unlock(ErlNifMutex*);

#define ERL_NIF_INIT(NAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD)                                   \
  int nxs_stub_nif_init_##LOAD(ErlNifEnv* env) {                                                    \
    (void)sizeof(FUNCS);                                                                           \
    return LOAD(env, NULL, (ERL_NIF_TERM)0);                                                       \
  }
#endif
