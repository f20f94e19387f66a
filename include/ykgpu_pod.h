/*
 * ykgpu_pod.h -- C ABI of the request-vector builder (part of libykgpu.so, host code, no GPU needed).
 *
 * The "vectors in" half of the boundary: turns what Kubernetes says about a pod's
 * containers, or a node's allocatable list, into the int64 resource vector the engine sweeps (yk_asks_upsert `req`,
 * yk_nodes_upsert `total`).  It replaces, with identical results:
 *   common.GetPodResource            /root/reference/pkg/common/resource.go:56-109
 *     computeContainerResource       :111-127   max(spec requests, status.allocatedResources, status.resources.requests),
 *                                               or status.resources.requests alone when a resize was found infeasible
 *     isResizeInfeasible             :132-142
 *     updateMax                      :145-162
 *     checkInitContainerRequest      :164-182   native sidecars (restartPolicy Always) accumulate, plain init containers peak
 *     getPodLevelResource            :287-301   pod-level requests override cpu / memory (/ hugepages-*) only
 *     pod.Spec.Overhead              :95-106    added on top
 *   common.GetNodeResource           :188-195   node.Status.Allocatable as is
 *   getResource                      :273-285   "cpu" becomes "vcore" in milli-units (Quantity.MilliValue), every other
 *                                               resource keeps its name and Quantity.Value
 * and the text form of quantities [EXT k8s.io/apimachinery pkg/api/resource, pinned by go.mod]: decimal numbers with
 * binary (Ki Mi Gi Ti Pi Ei), decimal (n u m k M G T P E) or exponent (e3 / E-2) suffixes; Value() and MilliValue()
 * round away from zero.
 * Pinned by the reference's own tables, transcribed into tests/golden/pod_resources.json:
 *   TestParsePodResource, TestInitContainerPodResources, TestGetPodResourcesWithPodLevelRequests,
 *   TestGetPodResourcesWithInPlacePodVerticalScaling, TestBestEffortPod, TestGPUOnlyResources, TestNodeResource
 *   (/root/reference/pkg/common/resource_test.go:153-840).
 *
 * All functions return YK_OK (0) or a negative yk_status of ykgpu.h; none of them touches CUDA.
 */
#ifndef YKGPU_POD_H
#define YKGPU_POD_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* resource.ParseQuantity(text): *value = Quantity.Value(), *milli = Quantity.MilliValue() (either may be NULL).
   YK_ERR_ARG on text the k8s parser rejects; an amount that does not fit int64 saturates at INT64_MAX. */
int yk_quantity_parse(const char* text, int64_t* value, int64_t* milli);

typedef struct yk_podspec yk_podspec;

yk_podspec* yk_podspec_new(void);
void yk_podspec_free(yk_podspec* p);
void yk_podspec_clear(yk_podspec* p);   /* reuse the object for the next pod */

/* container kinds, in the order they appear in pod.Spec.Containers / pod.Spec.InitContainers */
#define YK_CONTAINER 0u
#define YK_INIT_CONTAINER 1u
#define YK_INIT_SIDECAR 2u            /* init container with restartPolicy: Always (native sidecar) */
/* -> index of the new container (>= 0), or a negative status */
int yk_podspec_container(yk_podspec* p, const char* name, uint32_t kind);
/* pod.Status.ContainerStatuses / InitContainerStatuses has an entry with this name; has_resources = its .Resources != nil */
int yk_podspec_status(yk_podspec* p, const char* container_name, int32_t has_resources);

/* which resource list a quantity belongs to */
#define YK_LIST_REQUESTS 0u           /* container.Resources.Requests                (container = index) */
#define YK_LIST_ALLOCATED 1u          /* containerStatus.AllocatedResources          (container = index; needs yk_podspec_status) */
#define YK_LIST_STATUS_REQUESTS 2u    /* containerStatus.Resources.Requests          (container = index; needs has_resources) */
#define YK_LIST_POD_REQUESTS 3u       /* pod.Spec.Resources.Requests (pod level)     (container ignored) */
#define YK_LIST_OVERHEAD 4u           /* pod.Spec.Overhead                           (container ignored) */
int yk_podspec_quantity(yk_podspec* p, uint32_t list, int32_t container, const char* resource, const char* quantity);
/* same with an already parsed quantity (Value() and MilliValue()) */
int yk_podspec_amount(yk_podspec* p, uint32_t list, int32_t container, const char* resource, int64_t value, int64_t milli);

/* pod.Status.Resize == Infeasible, or a PodResizePending condition with reason Infeasible */
int yk_podspec_resize_infeasible(yk_podspec* p, int32_t yes);

/* GetPodResource: -> number of resources in the result ("pods" = 1 is always there), or a negative status */
int yk_podspec_compute(yk_podspec* p);
/* result entry i, ordered by resource name; names use the scheduler's vocabulary ("vcore", "memory", "pods", ...) */
int yk_podspec_result(const yk_podspec* p, uint32_t i, const char** name, int64_t* value);
/* the ABI vector: out[k] = result[dim_names[k]] or 0; *n_unmapped = result resources that no dimension names
   (such a pod must go the slow path: the engine would not account for that resource) */
int yk_podspec_vector(const yk_podspec* p, uint32_t n_dims, const char* const* dim_names, int64_t* out, uint32_t* n_unmapped);

/* GetNodeResource for one allocatable entry: *name_out = scheduler name of the resource ("cpu" -> "vcore"),
   *amount = milli-units for cpu, Value() otherwise.  name_out points into static or caller storage. */
int yk_node_quantity(const char* resource, const char* quantity, const char** name_out, int64_t* amount);

#ifdef __cplusplus
}
#endif
#endif
