/* TEST INFRASTRUCTURE -- stand-in for <Windows.h>: kernel.cu only uses GetTickCount (kernel.cu:1100). */
#ifndef DOGERAY_ORACLE_STUB_WINDOWS_H
#define DOGERAY_ORACLE_STUB_WINDOWS_H
#include <ctime>
#include <cstring>
static inline unsigned long GetTickCount() { return (unsigned long)(clock() / (CLOCKS_PER_SEC / 1000)); }
#endif
