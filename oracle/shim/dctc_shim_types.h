/* Type shim so that the reference's own src/dct.c and src/render.c compile UNMODIFIED without
 * glib / gtk / libgimp / liblqr (none of which exist in this image).
 *
 * TEST INFRASTRUCTURE ONLY: used by oracle/Makefile to build oracle/_ref/libdctc_ref.so from the
 * sources where they lie under /root/reference.  Nothing here is reference code: it only declares
 * the scalar typedefs, macros and opaque types those two files mention.  Only the hot-path
 * functions (render.c:122-157, dct.c:56-110) are ever called; every other gimp/gtk/lqr symbol stays
 * unresolved (-Wl,--unresolved-symbols=ignore-all) and is never reached.
 */
#ifndef DCTC_SHIM_TYPES_H
#define DCTC_SHIM_TYPES_H

#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

typedef int gint;
typedef unsigned int guint;
typedef int gint32;
typedef unsigned char guchar;
typedef unsigned char guint8;
typedef char gchar;
typedef float gfloat;
typedef double gdouble;
typedef int gboolean;
typedef void *gpointer;

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef ABS
#define ABS(a) (((a) < 0) ? -(a) : (a))
#endif
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef CLAMP
#define CLAMP(x, lo, hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x)))
#endif
#ifndef ROUND
#define ROUND(x) ((int) ((x) + 0.5))
#endif

#define g_new(type, n) ((type *) malloc(sizeof(type) * (size_t) (n)))
#define g_try_new(type, n) ((type *) malloc(sizeof(type) * (size_t) (n)))
#define g_free(p) free(p)
#define g_snprintf snprintf

/* gtk / gimp opaque handles */
typedef struct ShimGtkObject_ GtkObject;
typedef struct ShimGimpPreview_ GimpPreview;
typedef struct ShimGimpDrawable_ {
    gint32 drawable_id;
    guint width;
    guint height;
    guint bpp;
} GimpDrawable;
typedef struct ShimGimpPixelRgn_ {
    void *opaque[16];
} GimpPixelRgn;

enum { GIMP_RGB = 0, GIMP_GRAY = 1 };
enum { GIMP_RGB_IMAGE = 0, GIMP_GRAY_IMAGE = 2 };
enum { GIMP_NORMAL_MODE = 0 };
#define GIMP_DRAWABLE_PREVIEW(p) ((void *) (p))

GimpDrawable *gimp_drawable_get(gint32 id);
const gchar *gimp_drawable_get_name(gint32 id);
gboolean gimp_progress_init(const gchar *message);
gboolean gimp_progress_update(gdouble percentage);
gboolean gimp_progress_end(void);

/* liblqr opaque handles and the few prototypes render.c needs with non-int return types */
typedef struct ShimLqrCarver_ LqrCarver;
typedef struct ShimLqrProgress_ LqrProgress;
typedef struct ShimLqrVMap_ LqrVMap;
typedef struct ShimLqrVMapList_ LqrVMapList;
typedef struct ShimLqrReadingWindow_ LqrReadingWindow;
typedef int LqrRetVal;
typedef LqrRetVal (*LqrProgressFuncInit)(const gchar *);
typedef LqrRetVal (*LqrProgressFuncUpdate)(gdouble);
typedef LqrRetVal (*LqrProgressFuncEnd)(void);
typedef gfloat (*LqrEnergyFunc)(gint x, gint y, gint w, gint h, LqrReadingWindow *rw, gpointer extra);

enum { LQR_ER_BRIGHTNESS = 0, LQR_ER_LUMA = 1, LQR_ER_RGBA = 2, LQR_ER_CUSTOM = 3 };
enum { LQR_COLDEPTH_8I = 0 };
enum { LQR_GREY_IMAGE = 1 };
#define LQR_MAX_NAME_LENGTH 1024

gdouble lqr_rwindow_read(LqrReadingWindow *rw, gint x, gint y, gint channel);
gint lqr_rwindow_get_radius(LqrReadingWindow *rw);
LqrCarver *lqr_carver_new(guchar *buffer, gint width, gint height, gint channels);
LqrProgress *lqr_progress_new(void);
LqrVMapList *lqr_vmap_list_start(LqrCarver *r);
LqrVMap *lqr_vmap_list_current(LqrVMapList *list);
gint *lqr_vmap_get_data(LqrVMap *vmap);

#endif
