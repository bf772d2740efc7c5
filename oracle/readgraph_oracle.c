/* TEST INFRASTRUCTURE (oracle): CPU restatement of Assembler::createReadGraph, ReadGraph.creationMethod 0
 * (src/AssemblerReadGraph.cpp:35-175). Only tests/ may call it. Parity status: a restatement only (the member needs the
 * whole Assembler, which is unbuildable here); it follows the reference's control flow with qsort in place of
 * std::nth_element — the SET nth_element leaves in the first maxAlignmentCount places is the same. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t markerCount, alignmentId; } ReadAlignment;

/* std::greater<pair<uint32_t,uint32_t>>: descending by (markerCount, alignmentId) */
static int greaterFirst(const void* a, const void* b)
{
    const ReadAlignment* x = (const ReadAlignment*)a; const ReadAlignment* y = (const ReadAlignment*)b;
    if(x->markerCount != y->markerCount) return x->markerCount > y->markerCount ? -1 : 1;
    if(x->alignmentId != y->alignmentId) return x->alignmentId > y->alignmentId ? -1 : 1;
    return 0;
}

/* alignmentData: n 64-byte records (in/out: isInReadGraph). keep: uint8[n]. edges: room for 2n 16-byte records.
 * connToc: uint32[2*readCount+1], connData: uint32[4n]. Returns the number of edges. */
uint64_t orc_create_read_graph(uint32_t* alignmentData, uint64_t n, uint64_t readCount, uint32_t maxAlignmentCount,
                               uint8_t* keep, uint32_t* edges, uint32_t* connToc, uint32_t* connData)
{
    memset(keep, 0, n);
    /* the rows (readId, 0) of the alignment table: every alignment of the read (src/AssemblerAlign.cpp:509-571) */
    uint64_t* count = (uint64_t*)calloc(readCount + 1, sizeof(uint64_t));
    for(uint64_t a = 0; a < n; a++) { count[alignmentData[16 * a] + 1]++; count[alignmentData[16 * a + 1] + 1]++; }
    for(uint64_t r = 0; r < readCount; r++) count[r + 1] += count[r];
    ReadAlignment* all = (ReadAlignment*)malloc((2 * n + 1) * sizeof(ReadAlignment));
    uint64_t* fill = (uint64_t*)malloc((readCount + 1) * sizeof(uint64_t));
    memcpy(fill, count, (readCount + 1) * sizeof(uint64_t));
    for(uint64_t a = 0; a < n; a++) {
        for(int side = 0; side < 2; side++) {
            ReadAlignment* p = all + fill[alignmentData[16 * a + side]]++;
            p->markerCount = alignmentData[16 * a + 9];         /* AlignmentInfo::markerCount */
            p->alignmentId = (uint32_t)a;
        }
    }
    /* :53-86 */
    for(uint64_t r = 0; r < readCount; r++) {
        ReadAlignment* row = all + count[r];
        uint64_t size = count[r + 1] - count[r];
        if(size > maxAlignmentCount) {
            qsort(row, size, sizeof(ReadAlignment), greaterFirst);
            size = maxAlignmentCount;
        }
        for(uint64_t k = 0; k < size; k++) keep[row[k].alignmentId] = 1;
    }
    free(all); free(fill); free(count);
    /* createReadGraphUsingSelectedAlignments :95-144 */
    uint64_t edgeCount = 0;
    for(uint64_t a = 0; a < n; a++) {
        uint32_t* rec = alignmentData + 16 * a;
        rec[15] = (rec[15] & ~1u) | keep[a];
        if(!keep[a]) continue;
        const uint32_t o0 = 2 * rec[0], o1 = 2 * rec[1] + ((rec[2] & 0xff) ? 0u : 1u);
        for(uint32_t k = 0; k < 2; k++) {
            uint32_t* e = edges + 4 * edgeCount;
            e[0] = o0 ^ k; e[1] = o1 ^ k; e[2] = (uint32_t)a; e[3] = 0;
            edgeCount++;
        }
    }
    /* connectivity :147-159 */
    const uint64_t rows = 2 * readCount;
    memset(connToc, 0, (rows + 1) * sizeof(uint32_t));
    for(uint64_t i = 0; i < edgeCount; i++) { connToc[edges[4 * i] + 1]++; connToc[edges[4 * i + 1] + 1]++; }
    for(uint64_t r = 0; r < rows; r++) connToc[r + 1] += connToc[r];
    uint32_t* at = (uint32_t*)malloc((rows + 1) * sizeof(uint32_t));
    memcpy(at, connToc, (rows + 1) * sizeof(uint32_t));
    for(uint64_t i = 0; i < edgeCount; i++) {
        connData[at[edges[4 * i]]++] = (uint32_t)i;
        connData[at[edges[4 * i + 1]]++] = (uint32_t)i;
    }
    free(at);
    return edgeCount;
}
