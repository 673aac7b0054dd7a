/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A thin C-ABI driver around the UNMODIFIED reference library, compiled from
 * the sources where they lie under /root/reference (see oracle/Makefile).  It
 * replaces only the reference's file-reading front end and its resource
 * allocator, neither of which is on the factorize path:
 *
 *   - SparseFrame_read_matrix (CH.c:652-691) reads MatrixMarket text; here the
 *     triplets come straight from the caller and go through the reference's own
 *     SparseFrame_compress (CH.c:526-587) with the same workspace sizing rule
 *     (CH.c:679-684).
 *   - SparseFrame_allocate_gpu (CH.c:16-285) divides by zero without a GPU
 *     (CH.c:38) and derives devSlotSize from device memory (CH.c:199); here
 *     common_info is filled by hand: one CPU worker ("pretended GPU",
 *     CH.c:224-274), numGPU = 0, devSlotSize = caller's value.
 *
 * Everything else -- SparseFrame_analyze, SparseFrame_factorize,
 * SparseFrame_validate -- is the reference's own code, so the arrays this
 * driver hands back ARE the reference's results (CPU path, CH.c:2104-2148 /
 * LU.c:2622-2666).
 *
 * Built twice: -DSFREF_LU selects the LU headers/field set.
 */
#include "SparseFrame.h"

/* exported but undeclared entry points of the reference library */
int SparseFrame_initialize_matrix ( struct matrix_info_struct *matrix_info );
int SparseFrame_compress ( struct matrix_info_struct *matrix_info );
int SparseFrame_analyze ( struct common_info_struct *common_info, struct matrix_info_struct *matrix_info );
int SparseFrame_factorize ( struct common_info_struct *common_info, struct gpu_info_struct *gpu_info_list, struct matrix_info_struct *matrix_info );
int SparseFrame_validate ( struct matrix_info_struct *matrix_info );
int SparseFrame_cleanup_matrix ( struct matrix_info_struct *matrix_info );

struct sfref_handle
{
    struct common_info_struct common;
    struct gpu_info_struct gpu[1];
    struct matrix_info_struct matrix;
    int factorized;
    int validated;
};

/* is_complex: Tx holds nz interleaved (re, im) pairs -- the reference's Complex (Include/type.h), as its reader stores
 * a MatrixMarket file whose field is not "real" (CH.c:440-445, 505-511) */
static void *sfref_create_any ( int is_symmetric, int is_complex, long n, long nz, const long *Ti, const long *Tj, const double *Tx,
        unsigned long long dev_slot_size, int blas_threads )
{
    size_t elem = is_complex ? sizeof(Complex) : sizeof(Float);
    struct sfref_handle *h = calloc ( 1, sizeof ( struct sfref_handle ) );
    struct matrix_info_struct *m = &h->matrix;

    h->common.numCPU = 1;
    h->common.numGPU = 0;
    h->common.numGPU_physical = 0;
    h->common.matrixThreadNum = 1;
    h->common.numSparseMatrix = 1;
    h->common.devSlotSize = (size_t) dev_slot_size;

    memset ( &h->gpu[0], 0, sizeof ( h->gpu[0] ) );
    h->gpu[0].gpuIndex_physical = -1;
    h->gpu[0].lastMatrix = -1;
    omp_init_lock ( &h->gpu[0].gpuLock );

    if ( blas_threads > 0 )
        openblas_set_num_threads ( blas_threads );
    omp_set_nested ( TRUE );

    m->serial = 0;
    m->path = "<memory>";
    SparseFrame_initialize_matrix ( m );

    m->isSymmetric = is_symmetric;
    m->isComplex = is_complex ? TRUE : FALSE;
    m->ncol = n;
    m->nrow = n;
    m->nzmax = nz;
    m->Tj = malloc ( ( nz > 0 ? nz : 1 ) * sizeof(Long) );
    m->Ti = malloc ( ( nz > 0 ? nz : 1 ) * sizeof(Long) );
    m->Tx = malloc ( ( nz > 0 ? nz : 1 ) * elem );
    memcpy ( m->Tj, Tj, nz * sizeof(Long) );
    memcpy ( m->Ti, Ti, nz * sizeof(Long) );
    memcpy ( m->Tx, Tx, nz * elem );

    /* workspace sizing rule of SparseFrame_read_matrix; calloc because the LU
     * analysis counts into it without zeroing (LU.c:2271-2283) */
    m->workSize = MAX (
            ( 10 * m->nrow + ( 2 * m->nzmax - m->nrow ) + 1 ) * sizeof(Long),
            ( 3 * m->nrow + ( 2 * m->nzmax - m->nrow ) + 1 ) * sizeof(idx_t) );
    m->workspace = calloc ( 1, m->workSize );

    SparseFrame_compress ( m );

    return h;
}

void *sfref_create ( int is_symmetric, long n, long nz, const long *Ti, const long *Tj, const double *Tx,
        unsigned long long dev_slot_size, int blas_threads )
{
    return sfref_create_any ( is_symmetric, 0, n, nz, Ti, Tj, Tx, dev_slot_size, blas_threads );
}

/* The matrix comes from a MatrixMarket file through the reference's OWN reader (SparseFrame_read_matrix, CH.c:652-691:
 * getline + sscanf("%ld %ld %lg %lg"), then SparseFrame_compress) -- the oracle for the drop-in's reader.  The reference
 * ignores the reader's return code (CH.c:673); only well-formed files are defined input. */
int SparseFrame_read_matrix ( struct matrix_info_struct *matrix_info );

void *sfref_create_from_file ( const char *path, unsigned long long dev_slot_size, int blas_threads )
{
    struct sfref_handle *h = calloc ( 1, sizeof ( struct sfref_handle ) );
    struct matrix_info_struct *m = &h->matrix;

    h->common.numCPU = 1;
    h->common.numGPU = 0;
    h->common.numGPU_physical = 0;
    h->common.matrixThreadNum = 1;
    h->common.numSparseMatrix = 1;
    h->common.devSlotSize = (size_t) dev_slot_size;
    memset ( &h->gpu[0], 0, sizeof ( h->gpu[0] ) );
    h->gpu[0].gpuIndex_physical = -1;
    h->gpu[0].lastMatrix = -1;
    omp_init_lock ( &h->gpu[0].gpuLock );
    if ( blas_threads > 0 )
        openblas_set_num_threads ( blas_threads );
    omp_set_nested ( TRUE );

    m->serial = 0;
    m->path = strdup ( path );
    SparseFrame_initialize_matrix ( m );
    if ( SparseFrame_read_matrix ( m ) != 0 )
    {
        free ( h );
        return NULL;
    }
    return h;
}

void *sfref_create_complex ( int is_symmetric, long n, long nz, const long *Ti, const long *Tj, const double *Tz,
        unsigned long long dev_slot_size, int blas_threads )
{
    return sfref_create_any ( is_symmetric, 1, n, nz, Ti, Tj, Tz, dev_slot_size, blas_threads );
}

int sfref_analyze ( void *hv )
{
    struct sfref_handle *h = hv;
    return SparseFrame_analyze ( &h->common, &h->matrix );
}

int sfref_factorize ( void *hv )
{
    struct sfref_handle *h = hv;
    int rc = SparseFrame_factorize ( &h->common, h->gpu, &h->matrix );
    h->factorized = 1;
    return rc;
}

int sfref_validate ( void *hv )
{
    struct sfref_handle *h = hv;
    int rc;
    if ( h->validated )
    {
        free ( h->matrix.Bx ); free ( h->matrix.Xx ); free ( h->matrix.Rx );
        h->matrix.Bx = h->matrix.Xx = h->matrix.Rx = NULL;
    }
    rc = SparseFrame_validate ( &h->matrix );
    h->validated = 1;
    return rc;
}

long sfref_long ( void *hv, const char *name )
{
    struct matrix_info_struct *m = &( (struct sfref_handle*) hv )->matrix;
    if ( !strcmp ( name, "nrow" ) ) return m->nrow;
    if ( !strcmp ( name, "nzmax" ) ) return m->nzmax;
    if ( !strcmp ( name, "nsuper" ) ) return m->nsuper;
    if ( !strcmp ( name, "isize" ) ) return m->isize;
    if ( !strcmp ( name, "xsize" ) ) return m->xsize;
    if ( !strcmp ( name, "csize" ) ) return m->csize;
    if ( !strcmp ( name, "nsleaf" ) ) return m->nsleaf;
    if ( !strcmp ( name, "nstage" ) ) return m->nstage;
    if ( !strcmp ( name, "isSymmetric" ) ) return m->isSymmetric;
    if ( !strcmp ( name, "isComplex" ) ) return m->isComplex;
    return -1;
}

double sfref_double ( void *hv, const char *name )
{
    struct matrix_info_struct *m = &( (struct sfref_handle*) hv )->matrix;
    if ( !strcmp ( name, "analyzeTime" ) ) return m->analyzeTime;
    if ( !strcmp ( name, "factorizeTime" ) ) return m->factorizeTime;
    if ( !strcmp ( name, "solveTime" ) ) return m->solveTime;
    if ( !strcmp ( name, "residual" ) ) return m->residual;
    return -1;
}

/* pointer to a reference-owned array; valid until sfref_destroy */
const void *sfref_array ( void *hv, const char *name )
{
    struct matrix_info_struct *m = &( (struct sfref_handle*) hv )->matrix;
    if ( !strcmp ( name, "Cp" ) ) return m->Cp;
    if ( !strcmp ( name, "Ci" ) ) return m->Ci;
    if ( !strcmp ( name, "Cx" ) ) return m->Cx;
    if ( !strcmp ( name, "Lp" ) ) return m->Lp;
    if ( !strcmp ( name, "Li" ) ) return m->Li;
    if ( !strcmp ( name, "Lx" ) ) return m->Lx;
    if ( !strcmp ( name, "LTp" ) ) return m->LTp;
    if ( !strcmp ( name, "LTi" ) ) return m->LTi;
    if ( !strcmp ( name, "LTx" ) ) return m->LTx;
#ifdef SFREF_LU
    if ( !strcmp ( name, "Up" ) ) return m->Up;
    if ( !strcmp ( name, "Ui" ) ) return m->Ui;
    if ( !strcmp ( name, "Ux" ) ) return m->Ux;
    if ( !strcmp ( name, "UTp" ) ) return m->UTp;
    if ( !strcmp ( name, "UTi" ) ) return m->UTi;
    if ( !strcmp ( name, "UTx" ) ) return m->UTx;
#endif
    if ( !strcmp ( name, "Perm" ) ) return m->Perm;
    if ( !strcmp ( name, "Parent" ) ) return m->Parent;
    if ( !strcmp ( name, "Post" ) ) return m->Post;
    if ( !strcmp ( name, "ColCount" ) ) return m->ColCount;
    if ( !strcmp ( name, "Super" ) ) return m->Super;
    if ( !strcmp ( name, "SuperMap" ) ) return m->SuperMap;
    if ( !strcmp ( name, "Sparent" ) ) return m->Sparent;
    if ( !strcmp ( name, "LeafQueue" ) ) return m->LeafQueue;
    if ( !strcmp ( name, "Lsip" ) ) return m->Lsip;
    if ( !strcmp ( name, "Lsxp" ) ) return m->Lsxp;
    if ( !strcmp ( name, "Lsi" ) ) return m->Lsi;
    if ( !strcmp ( name, "Lsx" ) ) return m->Lsx;
    if ( !strcmp ( name, "ST_Map" ) ) return m->ST_Map;
    if ( !strcmp ( name, "ST_Pointer" ) ) return m->ST_Pointer;
    if ( !strcmp ( name, "ST_Index" ) ) return m->ST_Index;
    if ( !strcmp ( name, "Aoffset" ) ) return m->Aoffset;
    if ( !strcmp ( name, "Moffset" ) ) return m->Moffset;
    if ( !strcmp ( name, "Xx" ) ) return m->Xx;
    if ( !strcmp ( name, "Bx" ) ) return m->Bx;
    return NULL;
}

void sfref_destroy ( void *hv )
{
    struct sfref_handle *h = hv;
    SparseFrame_cleanup_matrix ( &h->matrix );
    omp_destroy_lock ( &h->gpu[0].gpuLock );
    free ( h );
}
