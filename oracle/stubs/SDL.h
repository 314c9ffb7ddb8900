/* TEST INFRASTRUCTURE -- headless stand-in for <SDL.h> so that the reference's
 * raygpu/kernel.cu compiles unmodified with nvcc for the GPU baseline
 * (oracle/make_ref.py).  Only the names kernel.cu mentions exist; the window
 * loop in its main() (renamed ref_main by -Dmain=ref_main) is never run. */
#ifndef DOGERAY_ORACLE_STUB_SDL_H
#define DOGERAY_ORACLE_STUB_SDL_H
#include <cstdio>
#include <cstdint>
typedef uint8_t Uint8;
typedef uint32_t Uint32;
struct SDL_Window {};
struct SDL_Renderer {};
struct SDL_Texture {};
struct SDL_Rect { int x, y, w, h; };
struct SDL_Surface { void* pixels; int pitch; };
struct SDL_Keysym { int sym; };
struct SDL_KeyboardEvent { SDL_Keysym keysym; };
struct SDL_Event { int type; SDL_KeyboardEvent key; };
enum { SDL_INIT_VIDEO = 0x20, SDL_MESSAGEBOX_ERROR = 0x10, SDL_PIXELFORMAT_ARGB8888 = 372645892,
       SDL_TEXTUREACCESS_STREAMING = 1, SDL_QUIT = 0x100, SDL_KEYDOWN = 0x300 };
enum { SDLK_ESCAPE = 27, SDLK_SPACE = 32, SDLK_b = 'b', SDLK_f = 'f', SDLK_g = 'g', SDLK_r = 'r', SDLK_s = 's',
       SDLK_t = 't', SDLK_w = 'w', SDLK_x = 'x', SDLK_z = 'z',
       SDLK_RIGHT = 1073741903, SDLK_LEFT, SDLK_DOWN, SDLK_UP,
       SDLK_KP_1 = 1073741913, SDLK_KP_2, SDLK_KP_3, SDLK_KP_4, SDLK_KP_5, SDLK_KP_6, SDLK_KP_7, SDLK_KP_8 };
static inline int SDL_Init(Uint32) { return -1; }
static inline const char* SDL_GetError() { return "headless stub"; }
static inline int SDL_ShowSimpleMessageBox(Uint32, const char* title, const char* msg, SDL_Window*) { fprintf(stderr, "[ref] %s: %s\n", title, msg); return 0; }
static inline int SDL_CreateWindowAndRenderer(int, int, Uint32, SDL_Window** w, SDL_Renderer** r) { *w = 0; *r = 0; return -1; }
static inline void SDL_SetWindowTitle(SDL_Window*, const char*) {}
static inline SDL_Texture* SDL_CreateTexture(SDL_Renderer*, Uint32, int, int, int) { return 0; }
static inline int SDL_LockTexture(SDL_Texture*, const SDL_Rect*, void**, int*) { return -1; }
static inline void SDL_UnlockTexture(SDL_Texture*) {}
static inline int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*) { return 0; }
static inline int SDL_SetRenderDrawColor(SDL_Renderer*, Uint8, Uint8, Uint8, Uint8) { return 0; }
static inline int SDL_RenderDrawPoint(SDL_Renderer*, int, int) { return 0; }
static inline void SDL_RenderPresent(SDL_Renderer*) {}
static inline int SDL_PollEvent(SDL_Event*) { return 0; }
static inline void SDL_DestroyRenderer(SDL_Renderer*) {}
static inline void SDL_DestroyWindow(SDL_Window*) {}
static inline void SDL_Quit() {}
static inline SDL_Surface* SDL_CreateRGBSurface(Uint32, int, int, int, Uint32, Uint32, Uint32, Uint32) { return 0; }
static inline int SDL_RenderReadPixels(SDL_Renderer*, const SDL_Rect*, Uint32, void*, int) { return -1; }
static inline int SDL_SaveBMP(SDL_Surface*, const char*) { return -1; }
static inline void SDL_FreeSurface(SDL_Surface*) {}
#endif
