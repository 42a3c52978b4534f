/* The part of the Open Image Denoise 2 C API that src/core/utility/denoise.c calls, declared from the library's public documentation
 * (TEST INFRASTRUCTURE: the implementation linked behind it is oracle/ref_host/fake_oidn.c, see its header). */
#pragma once
#include <stdbool.h>
#include <stddef.h>
typedef struct OIDNDeviceImpl* OIDNDevice;
typedef struct OIDNFilterImpl* OIDNFilter;
typedef struct OIDNBufferImpl* OIDNBuffer;
typedef enum { OIDN_DEVICE_TYPE_DEFAULT = 0, OIDN_DEVICE_TYPE_CPU = 1 } OIDNDeviceType;
typedef enum { OIDN_ERROR_NONE = 0, OIDN_ERROR_UNKNOWN = 1 } OIDNError;
typedef enum { OIDN_FORMAT_UNDEFINED = 0, OIDN_FORMAT_FLOAT = 1, OIDN_FORMAT_FLOAT2 = 2, OIDN_FORMAT_FLOAT3 = 3, OIDN_FORMAT_FLOAT4 = 4 } OIDNFormat;
typedef enum { OIDN_QUALITY_DEFAULT = 0, OIDN_QUALITY_FAST = 4, OIDN_QUALITY_BALANCED = 5, OIDN_QUALITY_HIGH = 6 } OIDNQuality;
OIDNDevice oidnNewDevice(OIDNDeviceType type);
void oidnCommitDevice(OIDNDevice device);
void oidnSyncDevice(OIDNDevice device);
void oidnReleaseDevice(OIDNDevice device);
OIDNError oidnGetDeviceError(OIDNDevice device, const char** outMessage);
OIDNFilter oidnNewFilter(OIDNDevice device, const char* type);
void oidnReleaseFilter(OIDNFilter filter);
void oidnSetFilterImage(OIDNFilter filter, const char* name, OIDNBuffer buffer, OIDNFormat format, size_t width, size_t height, size_t byteOffset, size_t pixelByteStride,
                        size_t rowByteStride);
void oidnSetFilterBool(OIDNFilter filter, const char* name, bool value);
void oidnSetFilterInt(OIDNFilter filter, const char* name, int value);
void oidnCommitFilter(OIDNFilter filter);
void oidnExecuteFilter(OIDNFilter filter);
OIDNBuffer oidnNewBuffer(OIDNDevice device, size_t byteSize);
void oidnReleaseBuffer(OIDNBuffer buffer);
void oidnWriteBuffer(OIDNBuffer buffer, size_t byteOffset, size_t byteSize, const void* srcHostPtr);
void oidnReadBuffer(OIDNBuffer buffer, size_t byteOffset, size_t byteSize, void* dstHostPtr);
