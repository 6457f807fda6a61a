/*
 * DenovoB200.java — Panama FFM (java.lang.foreign, JDK 22+) binding of include/denovo_b200.h.
 *
 * SOURCE ONLY: this image has no JDK (no javac, no JVM), so this file has never been compiled
 * here.  It is the binding a maintainer of alenzhao/denovo-variant-caller-java adds next to
 * DenovoCaller; the tested host is the C++ CLI (../../../../../../../host) over the same ABI.
 *
 * One downcall handle per entry point the caller plugins use (the context is a dnv_group: one dnv_ctx per
 * device, the first one also serves dnv_pileup / dnv_mendel_filter); struct layouts mirror the header byte for byte
 * (dnv_params 40 B, dnv_call 80 B, dnv_site_result 48 B, dnv_vcall 32 B, dnv_pcall 24 B).
 */
package com.google.cloud.genomics.denovo.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;
import static java.lang.foreign.ValueLayout.JAVA_SHORT;

public final class DenovoB200 implements AutoCloseable {

  public static final int OK = 0, E_INVALID = -1, E_CUDA = -2, E_CAPACITY = -3, E_BADBASE = -4,
      E_UNSORTED = -5, E_NOMEM = -6;
  public static final int METHOD_MAP = 0, METHOD_BAYES = 1, METHOD_LRT = 2; // InferenceMethod ordinals

  /** dnv_params */
  public static final StructLayout PARAMS = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("seq_err_rate"), JAVA_DOUBLE.withName("denovo_mut_rate"),
      JAVA_DOUBLE.withName("lrt_threshold"), JAVA_INT.withName("method"), JAVA_INT.withName("device"),
      JAVA_INT.withName("flags"), JAVA_INT.withName("reserved"));
  /** dnv_call: one called site (BayesCallResult + BayesInferenceResult) */
  public static final StructLayout CALL = MemoryLayout.structLayout(
      JAVA_LONG.withName("site"), MemoryLayout.sequenceLayout(12, JAVA_SHORT).withName("counts"),
      JAVA_BYTE.withName("dad"), JAVA_BYTE.withName("mom"), JAVA_BYTE.withName("child"),
      JAVA_BYTE.withName("is_denovo"), JAVA_INT.withName("flags"), JAVA_DOUBLE.withName("max_ll"),
      JAVA_DOUBLE.withName("p"), JAVA_DOUBLE.withName("lr"), JAVA_DOUBLE.withName("log_men"),
      JAVA_DOUBLE.withName("log_den"));
  /** dnv_vcall: one (Variant, VariantCall) pair of the VARIANT-mode stream */
  public static final StructLayout VCALL = MemoryLayout.structLayout(
      JAVA_LONG.withName("start"), JAVA_LONG.withName("end"), JAVA_INT.withName("variant_id"),
      JAVA_BYTE.withName("gt0"), JAVA_BYTE.withName("gt1"), JAVA_BYTE.withName("ngt"),
      JAVA_BYTE.withName("person"), JAVA_BYTE.withName("filter_pass"), JAVA_BYTE.withName("ref_len"),
      JAVA_BYTE.withName("ref_base"), JAVA_BYTE.withName("alt0_len"), JAVA_BYTE.withName("alt0_base"),
      JAVA_BYTE.withName("alt1_len"), JAVA_BYTE.withName("alt1_base"), JAVA_BYTE.withName("pad"));
  /** dnv_pcall: VariantsBuffer.PositionCall */
  public static final StructLayout PCALL = MemoryLayout.structLayout(
      JAVA_LONG.withName("position"), JAVA_LONG.withName("child_index"), JAVA_BYTE.withName("child"),
      JAVA_BYTE.withName("mom"), JAVA_BYTE.withName("dad"), JAVA_BYTE.withName("is_denovo"),
      MemoryLayout.paddingLayout(4));

  private static final Linker LINKER = Linker.nativeLinker();
  private final Arena libArena = Arena.ofShared();
  private final SymbolLookup lib;
  private final MethodHandle create, destroy, lastError, inferCounts, inferCounts8, pileup, mendelFilter, synthCountsAt;
  private final MethodHandle groupCreate, groupDestroy, groupInferCounts, groupInferCounts8, groupCtx, allocPinned, freePinned;
  private MemorySegment ctx = MemorySegment.NULL;    // dnv_ctx of the first device (pileup, Mendelian filter)
  private MemorySegment group = MemorySegment.NULL;  // dnv_group over all devices (READ-mode inference), owns ctx

  private MethodHandle handle(String name, FunctionDescriptor fd) {
    return LINKER.downcallHandle(lib.find(name).orElseThrow(
        () -> new UnsatisfiedLinkError("libdenovo_b200.so does not export " + name)), fd);
  }

  /**
   * Replaces {@code new BayesInfer(shared)} (ReadCaller.java:49-52): builds the CPTs once on the device.
   * There is no CPU fallback: without a CUDA device this throws.
   */
  public DenovoB200(String libraryPath, double seqErrRate, double denovoMutRate, double lrtThreshold,
      int method, int device) {
    this(libraryPath, seqErrRate, denovoMutRate, lrtThreshold, method, new int[] {device});
  }

  /**
   * Several devices driven by this one JVM (dnv_group): READ-mode batches are split by genomic interval over
   * them and the call records come back gathered in input order — the batching and result-gather layer of the
   * Java host is two downcalls.
   */
  public DenovoB200(String libraryPath, double seqErrRate, double denovoMutRate, double lrtThreshold,
      int method, int[] devices) {
    lib = SymbolLookup.libraryLookup(libraryPath, libArena);
    groupCreate = handle("dnv_group_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    groupDestroy = handle("dnv_group_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    groupInferCounts = handle("dnv_group_infer_counts", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
        ADDRESS, JAVA_LONG, ADDRESS));
    groupInferCounts8 = handle("dnv_group_infer_counts8", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
        ADDRESS, JAVA_LONG, ADDRESS));
    groupCtx = handle("dnv_group_ctx", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    allocPinned = handle("dnv_alloc_pinned", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, ADDRESS));
    freePinned = handle("dnv_free_pinned", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    create = handle("dnv_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    destroy = handle("dnv_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    lastError = handle("dnv_last_error", FunctionDescriptor.of(ADDRESS));
    inferCounts = handle("dnv_infer_counts", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
        ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    inferCounts8 = handle("dnv_infer_counts8", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
        ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    pileup = handle("dnv_pileup", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
        JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS));
    mendelFilter = handle("dnv_mendel_filter", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
        ADDRESS, JAVA_LONG, ADDRESS));
    synthCountsAt = handle("dnv_synth_counts_at", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS,
        JAVA_LONG, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    try (Arena a = Arena.ofConfined()) {
      MemorySegment p = a.allocate(PARAMS);
      p.set(JAVA_DOUBLE, 0, seqErrRate);
      p.set(JAVA_DOUBLE, 8, denovoMutRate);
      p.set(JAVA_DOUBLE, 16, lrtThreshold);
      p.set(JAVA_INT, 24, method);
      p.set(JAVA_INT, 28, devices[0]);
      p.set(JAVA_INT, 32, 0);
      MemorySegment dev = a.allocate((long) devices.length * 4, 4);
      for (int i = 0; i < devices.length; i++) dev.setAtIndex(JAVA_INT, i, devices[i]);
      MemorySegment out = a.allocate(ADDRESS);
      check((int) groupCreate.invoke(p, dev, devices.length, out));
      group = out.get(ADDRESS, 0);
      check((int) groupCtx.invoke(group, 0, out)); // borrowed: the group owns its contexts
      ctx = out.get(ADDRESS, 0);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  private void check(int rc) throws Throwable {
    if (rc == OK) return;
    MemorySegment msg = ((MemorySegment) lastError.invoke()).reinterpret(4096);
    String text = msg.getString(0);
    switch (rc) {
      case E_BADBASE: throw new IllegalArgumentException(text); // what Allele.valueOf throws
      case E_INVALID: throw new IllegalArgumentException(text);
      default: throw new IllegalStateException("denovo_b200 error " + rc + ": " + text);
    }
  }

  /**
   * Batched BayesInfer.infer (BayesInfer.java:50-69).  counts = uint16 [n][3][4], person order DAD, MOM,
   * CHILD, allele order A,C,G,T.  Returns the number of called sites written to calls (input order).
   */
  public long inferCounts(MemorySegment counts, long n, MemorySegment calls, long cap) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment nCalls = a.allocate(JAVA_LONG);
      check((int) inferCounts.invoke(ctx, counts, n, MemorySegment.NULL, calls, cap, nCalls));
      return nCalls.get(JAVA_LONG, 0);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /**
   * Same with counts packed as uint8 [n][3][4] (12 bytes per site): for pileups whose per-allele depth is
   * below 256 this halves the host-to-device traffic, which is what bounds the host path (PCIe).
   */
  public long inferCounts8(MemorySegment counts8, long n, MemorySegment calls, long cap) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment nCalls = a.allocate(JAVA_LONG);
      check((int) inferCounts8.invoke(ctx, counts8, n, MemorySegment.NULL, calls, cap, nCalls));
      return nCalls.get(JAVA_LONG, 0);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /**
   * inferCounts over all devices of the group: counts are split by interval (one host thread per device inside the
   * library), every device infers its shard, the records are gathered through the first device and returned in
   * input order.  Pass buffers from allocPinned so that the devices copy concurrently at the PCIe rate.
   */
  public long groupInferCounts(MemorySegment counts, long n, MemorySegment calls, long cap) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment nCalls = a.allocate(JAVA_LONG);
      check((int) groupInferCounts.invoke(group, counts, n, calls, cap, nCalls));
      return nCalls.get(JAVA_LONG, 0);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /** Page-locked host memory (dnv_alloc_pinned); release with freePinned. */
  public MemorySegment allocPinned(long bytes) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment out = a.allocate(ADDRESS);
      check((int) allocPinned.invoke(bytes, out));
      return out.get(ADDRESS, 0).reinterpret(bytes);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  public void freePinned(MemorySegment p) {
    try {
      freePinned.invoke(p);
    } catch (Throwable ignored) {
      // nothing useful to do
    }
  }

  /** Batched new ReadSummary(reads, position) x3 (ReadSummary.java:45-59). */
  public void pileup(MemorySegment readStart, MemorySegment baseOff, MemorySegment bases, MemorySegment person,
      long nReads, MemorySegment sitePos, long nSites, MemorySegment countsOut) {
    try {
      check((int) pileup.invoke(ctx, readStart, baseOff, bases, person, nReads, sitePos, nSites, countsOut));
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /** One shard of VariantCaller.callSimpleDenovo (VariantCaller.java:120-210). */
  public long mendelFilter(MemorySegment vcalls, long n, MemorySegment out, long cap) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment nOut = a.allocate(JAVA_LONG);
      check((int) mendelFilter.invoke(ctx, vcalls, n, out, cap, nOut));
      return nOut.get(JAVA_LONG, 0);
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /** Synthetic pileup source standing in for reads().search() (ReadCaller.java:155-186). */
  public void synthCountsAt(long seed, MemorySegment index, long n, MemorySegment thr, int nThr, int errThr,
      MemorySegment counts) {
    try {
      check((int) synthCountsAt.invoke(ctx, seed, index, n, thr, nThr, errThr, counts));
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  @Override
  public void close() {
    try {
      if (!group.equals(MemorySegment.NULL)) groupDestroy.invoke(group); // destroys its contexts
    } catch (Throwable ignored) {
      // nothing useful to do at shutdown
    }
    ctx = MemorySegment.NULL;
    group = MemorySegment.NULL;
    libArena.close();
  }
}
