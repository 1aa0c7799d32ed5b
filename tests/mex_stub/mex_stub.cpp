// Functional stand-in for the subset of MATLAB's MEX / mx API that matlab/manisdp_mex.cpp uses, plus a small C harness
// (stub_*) so that tests can build mxArrays, call mexFunction and read the outputs back WITHOUT MATLAB:
// tests/test_mex_gateway.py links this file with the unmodified gateway source and libmanisdp_b200.so and drives
// create -> set_Y -> tr_solve -> kkt -> get_Y on the GPU through it.  Test infrastructure only (never shipped).
//
// Semantics mirrored from the MEX API: column-major double matrices, CSC sparse matrices with mwIndex (uint64) ir/jc,
// 1 x 1 struct arrays of named fields, uint64 scalars, char row vectors; mexErrMsgIdAndTxt does not return (here: it
// throws, and stub_call turns the exception into an error code + message, as MATLAB turns it into an MException).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "mex.h"

struct mxArray_tag {
  enum Kind { DENSE, SPARSE, CHARS, STRUCT, UINT64, CELL } kind = DENSE;  // CELL: `fields` holds the elements
  size_t m = 0, n = 0;
  std::vector<double> pr;
  std::vector<mwIndex> ir, jc;
  std::string str;
  std::vector<std::string> names;
  std::vector<mxArray*> fields;
  std::vector<uint64_t> u64;
  ~mxArray_tag() {
    for (mxArray* f : fields) delete f;
  }
};

struct MexError : std::runtime_error {
  std::string id;
  MexError(const std::string& i, const std::string& msg) : std::runtime_error(msg), id(i) {}
};

static bool g_locked = false;
static void (*g_at_exit)(void) = nullptr;

extern "C" {
bool mxIsChar(const mxArray* a) { return a && a->kind == mxArray::CHARS; }
bool mxIsSparse(const mxArray* a) { return a && a->kind == mxArray::SPARSE; }
bool mxIsDouble(const mxArray* a) { return a && (a->kind == mxArray::DENSE || a->kind == mxArray::SPARSE); }
bool mxIsStruct(const mxArray* a) { return a && a->kind == mxArray::STRUCT; }
bool mxIsUint64(const mxArray* a) { return a && a->kind == mxArray::UINT64; }
bool mxIsCell(const mxArray* a) { return a && a->kind == mxArray::CELL; }
mxArray* mxCreateCellMatrix(mwSize m, mwSize n) {
  mxArray* a = new mxArray;
  a->kind = mxArray::CELL;
  a->m = m;
  a->n = n;
  a->fields.assign(m * n, nullptr);
  return a;
}
mxArray* mxGetCell(const mxArray* a, mwIndex i) { return a->fields.at((size_t)i); }
void mxSetCell(mxArray* a, mwIndex i, mxArray* v) {
  delete a->fields.at((size_t)i);
  a->fields[(size_t)i] = v;
}
int mxGetString(const mxArray* a, char* buf, mwSize len) {
  if (!mxIsChar(a) || len == 0) return 1;
  strncpy(buf, a->str.c_str(), len - 1);
  buf[len - 1] = 0;
  return a->str.size() >= len ? 1 : 0;
}
double mxGetScalar(const mxArray* a) {
  if (a->kind == mxArray::UINT64) return (double)a->u64.at(0);
  return a->pr.at(0);
}
double* mxGetPr(const mxArray* a) { return const_cast<double*>(a->pr.data()); }
void* mxGetData(const mxArray* a) {
  return a->kind == mxArray::UINT64 ? (void*)const_cast<uint64_t*>(a->u64.data()) : (void*)const_cast<double*>(a->pr.data());
}
mwIndex* mxGetIr(const mxArray* a) { return const_cast<mwIndex*>(a->ir.data()); }
mwIndex* mxGetJc(const mxArray* a) { return const_cast<mwIndex*>(a->jc.data()); }
size_t mxGetM(const mxArray* a) { return a->m; }
size_t mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
mxArray* mxGetField(const mxArray* a, mwIndex, const char* name) {
  for (size_t i = 0; i < a->names.size(); ++i)
    if (a->names[i] == name) return a->fields[i];
  return nullptr;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) {
  mxArray* a = new mxArray;
  a->kind = mxArray::DENSE;
  a->m = m;
  a->n = n;
  a->pr.assign(m * n, 0.0);
  return a;
}
mxArray* mxCreateDoubleScalar(double v) {
  mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
  a->pr[0] = v;
  return a;
}
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity) {
  if (cls != mxUINT64_CLASS) throw MexError("stub:class", "only uint64 numeric matrices are modelled");
  mxArray* a = new mxArray;
  a->kind = mxArray::UINT64;
  a->m = m;
  a->n = n;
  a->u64.assign(m * n, 0);
  return a;
}
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nf, const char** names) {
  mxArray* a = new mxArray;
  a->kind = mxArray::STRUCT;
  a->m = m;
  a->n = n;
  for (int i = 0; i < nf; ++i) {
    a->names.push_back(names[i]);
    a->fields.push_back(nullptr);
  }
  return a;
}
void mxSetFieldByNumber(mxArray* a, mwIndex, int i, mxArray* v) {
  delete a->fields.at((size_t)i);
  a->fields[(size_t)i] = v;
}
void mxDestroyArray(mxArray* a) { delete a; }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw MexError(id, buf);
}
void mexLock(void) { g_locked = true; }
int mexAtExit(void (*f)(void)) {
  g_at_exit = f;
  return 0;
}

// ---- harness ---------------------------------------------------------------------------------------------------------
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

mxArray* stub_string(const char* s) {
  mxArray* a = new mxArray;
  a->kind = mxArray::CHARS;
  a->m = 1;
  a->str = s;
  a->n = a->str.size();
  return a;
}
mxArray* stub_dense(size_t m, size_t n, const double* data) {  // column-major
  mxArray* a = mxCreateDoubleMatrix(m, n, mxREAL);
  if (data) memcpy(a->pr.data(), data, m * n * sizeof(double));
  return a;
}
mxArray* stub_sparse(size_t m, size_t n, const uint64_t* jc, const uint64_t* ir, const double* pr) {
  mxArray* a = new mxArray;
  a->kind = mxArray::SPARSE;
  a->m = m;
  a->n = n;
  a->jc.assign(jc, jc + n + 1);
  a->ir.assign(ir, ir + jc[n]);
  a->pr.assign(pr, pr + jc[n]);
  return a;
}
mxArray* stub_struct(int nf, const char** names, const double* vals) {
  mxArray* a = mxCreateStructMatrix(1, 1, nf, names);
  for (int i = 0; i < nf; ++i) mxSetFieldByNumber(a, 0, i, mxCreateDoubleScalar(vals[i]));
  return a;
}
mxArray* stub_cell(size_t n) { return mxCreateCellMatrix(n, 1); }
void stub_cell_set(mxArray* a, size_t i, mxArray* v) { mxSetCell(a, i, v); }  // the cell takes ownership of v
const mxArray* stub_cell_get(const mxArray* a, size_t i) { return mxGetCell(a, i); }
void stub_free(mxArray* a) { delete a; }
size_t stub_rows(const mxArray* a) { return a->m; }
size_t stub_cols(const mxArray* a) { return a->n; }
int stub_kind(const mxArray* a) { return (int)a->kind; }
const double* stub_data(const mxArray* a) { return a->pr.data(); }
uint64_t stub_u64(const mxArray* a) { return a->u64.at(0); }
int stub_field(const mxArray* a, const char* name, double* out) {
  const mxArray* f = mxGetField(a, 0, name);
  if (!f) return 1;
  *out = mxGetScalar(f);
  return 0;
}
// 0: ok; 1: the gateway raised mexErrMsgIdAndTxt (id / message copied out)
int stub_call(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs, char* id, char* msg, size_t cap) {
  try {
    mexFunction(nlhs, plhs, nrhs, prhs);
    return 0;
  } catch (const MexError& e) {
    if (id && cap) snprintf(id, cap, "%s", e.id.c_str());
    if (msg && cap) snprintf(msg, cap, "%s", e.what());
    return 1;
  }
}
int stub_is_locked(void) { return g_locked ? 1 : 0; }
void stub_run_at_exit(void) {
  if (g_at_exit) g_at_exit();
}
}
