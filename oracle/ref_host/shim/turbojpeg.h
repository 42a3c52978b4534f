/* Names of libjpeg-turbo's TurboJPEG API that src/core/utility/export/image.c mentions (TEST INFRASTRUCTURE; stubs, never called). */
#pragma once
typedef void* tjhandle;
enum { TJPF_RGBA = 7, TJSAMP_444 = 0, TJFLAG_ACCURATEDCT = 4096 };
tjhandle tjInitCompress(void);
int tjCompress2(tjhandle handle, const unsigned char* srcBuf, int width, int pitch, int height, int pixelFormat, unsigned char** jpegBuf, unsigned long* jpegSize, int jpegSubsamp,
                int jpegQual, int flags);
int tjDestroy(tjhandle handle);
void tjFree(unsigned char* buffer);
char* tjGetErrorStr(void);
char* tjGetErrorStr2(tjhandle handle);
