/*
 * tmjx_xla_ffi_c_api_min.h — the subset of XLA's FFI C API (`xla/ffi/api/c_api.h`, XLA FFI API version 0.1 as shipped with
 * jaxlib 0.4.31 .. 0.6.x) that the tmjx adapter touches, declared here ONLY because this image has no jaxlib.
 *
 * RECALLED, NOT COPIED: there is no XLA source or header in this environment; the declarations below were written from knowledge of
 * that header (struct members in declaration order, enum values).  `tmjx_xla_ffi.cc` includes the REAL header instead whenever it
 * is on the include path (`__has_include`), which makes the adapter binary-correct by construction; with these fallback
 * declarations it is compiled and exercised end to end against a hand-built call frame (tests/test_gpu_ffi.py), which checks the
 * adapter's own logic (operand / result / attribute decoding, dispatch, error path), not XLA's ABI.  A maintainer with jaxlib
 * should build with `-I $(python -c 'import jaxlib, os; print(os.path.dirname(jaxlib.__file__) + "/include")')`.
 */
#ifndef TMJX_XLA_FFI_C_API_MIN_H_
#define TMJX_XLA_FFI_C_API_MIN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XLA_FFI_API_MAJOR 0
#define XLA_FFI_API_MINOR 1

typedef struct XLA_FFI_Api XLA_FFI_Api;
typedef struct XLA_FFI_InternalApi XLA_FFI_InternalApi;
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_ExecutionContext XLA_FFI_ExecutionContext;
typedef struct XLA_FFI_Future XLA_FFI_Future;

typedef enum { XLA_FFI_Extension_Metadata = 1 } XLA_FFI_Extension_Type;
typedef struct XLA_FFI_Extension_Base {
  size_t struct_size;
  XLA_FFI_Extension_Type type;
  struct XLA_FFI_Extension_Base* next;
} XLA_FFI_Extension_Base;

typedef struct XLA_FFI_Api_Version {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int major_version;  /* out */
  int minor_version;  /* out */
} XLA_FFI_Api_Version;

typedef uint32_t XLA_FFI_Handler_Traits;
typedef struct XLA_FFI_Metadata {
  size_t struct_size;
  XLA_FFI_Api_Version api_version;
  XLA_FFI_Handler_Traits traits;
} XLA_FFI_Metadata;
typedef struct XLA_FFI_Metadata_Extension {
  XLA_FFI_Extension_Base extension_base;
  XLA_FFI_Metadata* metadata;
} XLA_FFI_Metadata_Extension;

typedef enum {
  XLA_FFI_Error_Code_OK = 0, XLA_FFI_Error_Code_CANCELLED = 1, XLA_FFI_Error_Code_UNKNOWN = 2, XLA_FFI_Error_Code_INVALID_ARGUMENT = 3,
  XLA_FFI_Error_Code_DEADLINE_EXCEEDED = 4, XLA_FFI_Error_Code_NOT_FOUND = 5, XLA_FFI_Error_Code_ALREADY_EXISTS = 6,
  XLA_FFI_Error_Code_PERMISSION_DENIED = 7, XLA_FFI_Error_Code_RESOURCE_EXHAUSTED = 8, XLA_FFI_Error_Code_FAILED_PRECONDITION = 9,
  XLA_FFI_Error_Code_ABORTED = 10, XLA_FFI_Error_Code_OUT_OF_RANGE = 11, XLA_FFI_Error_Code_UNIMPLEMENTED = 12,
  XLA_FFI_Error_Code_INTERNAL = 13, XLA_FFI_Error_Code_UNAVAILABLE = 14, XLA_FFI_Error_Code_DATA_LOSS = 15,
  XLA_FFI_Error_Code_UNAUTHENTICATED = 16
} XLA_FFI_Error_Code;

typedef struct XLA_FFI_Error_Create_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const char* message;
  XLA_FFI_Error_Code errc;
} XLA_FFI_Error_Create_Args;
typedef XLA_FFI_Error* XLA_FFI_Error_Create(XLA_FFI_Error_Create_Args* args);

typedef enum {
  XLA_FFI_DataType_INVALID = 0, XLA_FFI_DataType_PRED = 1, XLA_FFI_DataType_S8 = 2, XLA_FFI_DataType_S16 = 3, XLA_FFI_DataType_S32 = 4,
  XLA_FFI_DataType_S64 = 5, XLA_FFI_DataType_U8 = 6, XLA_FFI_DataType_U16 = 7, XLA_FFI_DataType_U32 = 8, XLA_FFI_DataType_U64 = 9,
  XLA_FFI_DataType_F16 = 10, XLA_FFI_DataType_F32 = 11, XLA_FFI_DataType_F64 = 12, XLA_FFI_DataType_BF16 = 16
} XLA_FFI_DataType;

typedef struct XLA_FFI_Buffer {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_DataType dtype;
  void* data;
  int64_t rank;
  int64_t* dims;  /* length == rank */
} XLA_FFI_Buffer;

typedef enum { XLA_FFI_ArgType_BUFFER = 1 } XLA_FFI_ArgType;
typedef enum { XLA_FFI_RetType_BUFFER = 1 } XLA_FFI_RetType;
typedef struct XLA_FFI_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_ArgType* types;  /* length == size */
  void** args;             /* length == size */
} XLA_FFI_Args;
typedef struct XLA_FFI_Rets {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_RetType* types;
  void** rets;
} XLA_FFI_Rets;

typedef struct XLA_FFI_ByteSpan { const char* ptr; size_t len; } XLA_FFI_ByteSpan;
typedef struct XLA_FFI_Scalar { XLA_FFI_DataType dtype; void* value; } XLA_FFI_Scalar;
typedef enum { XLA_FFI_AttrType_ARRAY = 1, XLA_FFI_AttrType_DICTIONARY = 2, XLA_FFI_AttrType_SCALAR = 3, XLA_FFI_AttrType_STRING = 4 } XLA_FFI_AttrType;
typedef struct XLA_FFI_Attrs {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_AttrType* types;
  XLA_FFI_ByteSpan** names;
  void** attr;
} XLA_FFI_Attrs;

typedef enum {
  XLA_FFI_ExecutionStage_INSTANTIATE = 0, XLA_FFI_ExecutionStage_PREPARE = 1, XLA_FFI_ExecutionStage_INITIALIZE = 2,
  XLA_FFI_ExecutionStage_EXECUTE = 3
} XLA_FFI_ExecutionStage;

typedef struct XLA_FFI_CallFrame {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const XLA_FFI_Api* api;
  XLA_FFI_ExecutionContext* ctx;
  XLA_FFI_ExecutionStage stage;
  XLA_FFI_Args args;
  XLA_FFI_Rets rets;
  XLA_FFI_Attrs attrs;
  XLA_FFI_Future* future;  /* out, optional */
} XLA_FFI_CallFrame;

typedef XLA_FFI_Error* XLA_FFI_Handler(XLA_FFI_CallFrame* call_frame);

typedef struct XLA_FFI_Stream_Get_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_ExecutionContext* ctx;
  void* stream;  /* out */
} XLA_FFI_Stream_Get_Args;
typedef XLA_FFI_Error* XLA_FFI_Stream_Get(XLA_FFI_Stream_Get_Args* args);

/* Only the leading members the adapter dereferences; the real struct continues with more function pointers. */
struct XLA_FFI_Api {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_Api_Version api_version;
  XLA_FFI_InternalApi* internal_api;
  XLA_FFI_Error_Create* XLA_FFI_Error_Create;
  void* XLA_FFI_Error_GetMessage;
  void* XLA_FFI_Error_Destroy;
  void* XLA_FFI_Handler_Register;
  XLA_FFI_Stream_Get* XLA_FFI_Stream_Get;
};

#ifdef __cplusplus
}
#endif
#endif /* TMJX_XLA_FFI_C_API_MIN_H_ */
