defmodule NxSignalB200.NIF do
  @moduledoc false
  # Loads priv/nxsignal_nif.so built from elixir/c_src/nxsignal_nif.c over include/nxsignal_b200.h.
  # Every stub below has a {name, arity} entry in the C file's funcs[] (tests/test_elixir_boundary.py
  # checks the two lists against each other; there is no BEAM in this repository's image).
  @on_load :load
  def load, do: :erlang.load_nif(~c"#{:code.priv_dir(:nx_signal_b200)}/nxsignal_nif", 0)
  def ctx_create(_dev), do: :erlang.nif_error(:not_loaded)
  def stft(_c, _x, _ch, _len, _w, _hop, _nfft, _pad, _lo, _hi, _scal, _sr), do: :erlang.nif_error(:not_loaded)
  def stft_c64(_c, _x, _ch, _len, _w, _hop, _nfft, _pad, _lo, _hi, _scal, _sr), do: :erlang.nif_error(:not_loaded)
  def istft(_c, _z, _ch, _frames, _zlen, _w, _hop, _nfft, _scal, _sr), do: :erlang.nif_error(:not_loaded)
  def istft_c2r(_c, _z, _ch, _frames, _zlen, _w, _hop, _nfft, _scal, _sr), do: :erlang.nif_error(:not_loaded)
  def fir(_c, _x, _ch, _len, _h, _mode), do: :erlang.nif_error(:not_loaded)
  def convolve_nd(_c, _a, _a_shape, _b, _b_shape, _complex, _mode), do: :erlang.nif_error(:not_loaded)
  def as_windowed(_c, _x, _elem, _ch, _len, _wl, _stride, _pad, _lo, _hi), do: :erlang.nif_error(:not_loaded)
  def overlap_and_add(_c, _t, _complex, _batch, _frames, _flen, _overlap), do: :erlang.nif_error(:not_loaded)
  def stft_to_mel(_c, _z, _ch, _frames, _zlen, _nfft, _mels, _sr, _max_mel, _f_sp), do: :erlang.nif_error(:not_loaded)

  def stft_mel(_c, _x, _ch, _len, _w, _hop, _nfft, _pad, _lo, _hi, _scal, _sr, _mels, _max_mel, _f_sp),
    do: :erlang.nif_error(:not_loaded)

  def median(_c, _t, _shape, _kernel_shape), do: :erlang.nif_error(:not_loaded)
  def wiener(_c, _t, _is_f64, _shape, _kernel_size, _has_noise, _noise), do: :erlang.nif_error(:not_loaded)
  def argrelextrema(_c, _data, _shape, _axis, _order, _comparator), do: :erlang.nif_error(:not_loaded)
  def window(_kind, _n, _periodic, _beta, _eps), do: :erlang.nif_error(:not_loaded)
  def firwin(_taps, _cutoffs, _window_kind, _beta, _pass_zero, _scale, _sr), do: :erlang.nif_error(:not_loaded)
  def fft_frequencies(_sr, _nfft), do: :erlang.nif_error(:not_loaded)
  def mel_filters(_nfft, _mels, _sr, _max_mel, _f_sp), do: :erlang.nif_error(:not_loaded)
end

defmodule NxSignalB200.Ctx do
  @moduledoc false
  # One nxs_ctx per device, created once and kept in :persistent_term.  Creation is serialised with
  # :global.trans so that two processes racing on the first call cannot both create (and leak) one;
  # calls through a context are serialised by the mutex its NIF resource carries, so any number
  # of BEAM processes may use it concurrently (dirty IO schedulers).
  def get(device \\ 0) do
    key = {__MODULE__, device}

    case :persistent_term.get(key, nil) do
      nil ->
        :global.trans({key, self()}, fn ->
          case :persistent_term.get(key, nil) do
            nil ->
              {:ok, c} = NxSignalB200.NIF.ctx_create(device)
              :persistent_term.put(key, c)
              c

            c ->
              c
          end
        end)

      c ->
        c
    end
  end

  def expr?(%Nx.Tensor{data: %Nx.Defn.Expr{}}), do: true
  def expr?(_), do: false

  def raise_nif({:error, :argument_error, msg}), do: raise(ArgumentError, List.to_string(msg))
  def raise_nif({:error, _, msg}), do: raise(RuntimeError, List.to_string(msg))

  # devectorised tensor, its vectorised axes, and the element count of everything but the last `keep` axes
  def flatten_batch(t, keep) do
    flat = Nx.devectorize(t)
    dims = Tuple.to_list(Nx.shape(flat))
    {lead, tail} = Enum.split(dims, length(dims) - keep)
    {flat, t.vectorized_axes, lead, tail, Enum.product(lead)}
  end
end

defmodule NxSignalB200 do
  @moduledoc """
  Drop-in heads for the accelerated path of `NxSignal`: the same names, arities, options, defaults,
  `ArgumentError` texts and return shapes as the reference, served by the B200 backend through
  `NxSignalB200.NIF`:

    * `stft/3` (`lib/nx_signal.ex:68`), `istft/3` (`:582`), `as_windowed/2` (`:249`),
      `overlap_and_add/2` (`:684`), `fft_frequencies/2` (`:154`), `mel_filters/4` (`:397`),
      `stft_to_mel/3` (`:486`);
    * `NxSignalB200.Windows.*` (`lib/nx_signal/windows.ex:33-341`);
    * `NxSignalB200.Filters.firwin/3`, `median/2`, `wiener/2` (`lib/nx_signal/filters.ex:17,80,147`);
    * `NxSignalB200.Convolution.convolve/3`, `correlate/3`, `fftconvolve/3`
      (`lib/nx_signal/convolution.ex:38,87,252`);
    * `NxSignalB200.PeakFinding.argrelmin/2`, `argrelmax/2`, `argrelextrema/3`
      (`lib/nx_signal/peak_finding.ex:131,252,340`);
    * extensions that are not reference heads: `stft_mel/3` (fused STFT -> log-mel) and `istft_c2r/3`.

  Called with `Nx.Defn.Expr` tensors (inside a user `defn`) every head falls back to `NxSignal`
  itself, so `alias NxSignalB200, as: NxSignal` at a call site changes nothing but the speed.
  Source-only here (no BEAM in this repository's image): `nx_signal_b200/__init__.py` mirrors these
  heads one for one and is the exercised binding.
  """

  alias NxSignalB200.{Ctx, NIF}

  @pad %{valid: 0, same: 1, reflect: 2}
  @scaling %{nil => 0, spectrum: 1, psd: 2}

  @doc """
  True when the NIF is loaded and a CUDA device answers (the backend has no CPU path: when this is false the
  reference's own graphs are the only implementation).  `NxSignal`'s heads can delegate on it:

      if NxSignalB200.available?() and not match?(%Nx.Tensor{data: %Nx.Defn.Expr{}}, data), do: NxSignalB200.stft(...)
  """
  def available? do
    try do
      is_reference(Ctx.get())
    rescue
      _ -> false
    catch
      _, _ -> false
    end
  end

  defp next_pow2(n), do: Bitwise.bsl(1, ceil(:math.log2(n)))

  defp fft_len(:power_of_two, n), do: next_pow2(n)
  defp fft_len(n, _), do: n

  # :valid | :same | :reflect | [{lo, hi}] -> {mode, lo, hi}; the reference's message otherwise (lib/nx_signal.ex:325-329)
  defp padding(p) when is_map_key(@pad, p), do: {@pad[p], 0, 0}
  defp padding([{lo, hi}]) when is_integer(lo) and is_integer(hi), do: {3, lo, hi}

  defp padding(config) when is_list(config),
    do:
      raise(
        ArgumentError,
        "padding must be a list of {high, low} tuples, where each element is an integer. Got: #{inspect(config)}"
      )

  defp padding(mode),
    do:
      raise(
        ArgumentError,
        "invalid padding mode specified, padding must be one of :valid, :same, or a padding configuration, got: #{inspect(mode)}"
      )

  defp scaling!(s) do
    case Map.fetch(@scaling, s) do
      {:ok, v} -> v
      :error -> raise ArgumentError, "invalid :scaling, expected one of :spectrum, :psd or nil, got: #{inspect(s)}"
    end
  end

  @doc "`NxSignal.stft/3` (lib/nx_signal.ex:68-130): `{z {frames, frequencies} c64, times, frequencies}`."
  def stft(data, window, opts \\ []) do
    if Ctx.expr?(data) or Ctx.expr?(window) do
      NxSignal.stft(data, window, opts)
    else
      {frame_length} = Nx.shape(window)

      opts =
        Keyword.validate!(opts, [
          :overlap_length,
          :window,
          :scaling,
          window_padding: :valid,
          sampling_rate: 100,
          fft_length: :power_of_two
        ])

      sampling_rate = opts[:sampling_rate] || raise ArgumentError, "missing sampling_rate option"
      overlap = opts[:overlap_length] || div(frame_length, 2)
      scaling = scaling!(opts[:scaling])
      {pad, lo, hi} = padding(opts[:window_padding])
      nfft = fft_len(opts[:fft_length], frame_length)
      {flat, vec_axes, lead, [len], ch} = Ctx.flatten_batch(data, 1)

      # complex data is a complex transform in the reference's graph (Nx.multiply -> Nx.fft, :101-102)
      {nif, wire} = if Nx.Type.complex?(Nx.type(flat)), do: {&NIF.stft_c64/12, :c64}, else: {&NIF.stft/12, :f32}

      case nif.(Ctx.get(), Nx.to_binary(Nx.as_type(flat, wire)), ch, len, Nx.to_binary(Nx.as_type(window, :f32)),
             frame_length - overlap, nfft, pad, lo, hi, scaling, sampling_rate * 1.0) do
        {:ok, z, times, freqs, frames} ->
          z =
            z
            |> Nx.from_binary(:c64)
            |> Nx.reshape(List.to_tuple(lead ++ [frames, nfft]))
            |> Nx.vectorize(vec_axes)
            |> then(&Nx.reshape(&1, &1.shape, names: [:frames, :frequencies]))

          {z, times |> Nx.from_binary(:f32) |> Nx.reshape({frames}, names: [:frames]),
           freqs |> Nx.from_binary(:f32) |> Nx.reshape({nfft}, names: [:frequencies])}

        err ->
          Ctx.raise_nif(err)
      end
    end
  end

  @doc "`NxSignal.istft/3` (lib/nx_signal.ex:582-638)."
  def istft(data, window, opts) do
    if Ctx.expr?(data) or Ctx.expr?(window) do
      NxSignal.istft(data, window, opts)
    else
      {y, _} = istft_impl(data, window, opts, &NIF.istft/10, :c64)
      y
    end
  end

  @doc """
  Opt-in counterpart of the one-sided STFT: `data {frames, fft_length/2 + 1}` c64 in, REAL signal out,
  `Re(istft(ext(z)))` of the reference head with `ext` the conjugate-mirror extension.  `:fft_length`
  is mandatory (it cannot be inferred from the one-sided bin count).
  """
  def istft_c2r(data, window, opts) do
    {y, _} = istft_impl(data, window, opts, &NIF.istft_c2r/10, :f32)
    y
  end

  defp istft_impl(data, window, opts, nif, out_type) do
    opts = Keyword.validate!(opts, [:fft_length, :overlap_length, :scaling, sampling_rate: 1000])
    n = Nx.size(window)
    {flat, vec_axes, lead, [frames, zlen], ch} = Ctx.flatten_batch(data, 2)
    nfft = opts[:fft_length] || next_pow2(zlen)
    overlap = opts[:overlap_length] || div(n, 2)

    if opts[:scaling] == :psd and is_nil(opts[:sampling_rate]),
      do: raise(ArgumentError, ":sampling_rate is mandatory if scaling is :psd")

    scaling = scaling!(opts[:scaling])

    if overlap >= n,
      do: raise(ArgumentError, "overlap_length must be a number less than the window size #{n}, got: #{inspect(n)}")

    case nif.(Ctx.get(), Nx.to_binary(Nx.as_type(flat, :c64)), ch, frames, zlen, Nx.to_binary(Nx.as_type(window, :f32)),
           n - overlap, nfft, scaling, (opts[:sampling_rate] || 1000) * 1.0) do
      {:ok, y, out_len} ->
        {y |> Nx.from_binary(out_type) |> Nx.reshape(List.to_tuple(lead ++ [out_len])) |> Nx.vectorize(vec_axes), out_len}

      err ->
        Ctx.raise_nif(err)
    end
  end

  @doc "`NxSignal.as_windowed/2` (lib/nx_signal.ex:249-364): `{num_windows, window_length}` per vectorised entry."
  def as_windowed(tensor, opts \\ []) do
    if Ctx.expr?(tensor) do
      NxSignal.as_windowed(tensor, opts)
    else
      opts = Keyword.validate!(opts, [:window_length, padding: :valid, stride: 1])
      window_length = opts[:window_length]

      stride =
        case opts[:stride] do
          [s] when is_integer(s) and s >= 1 -> s
          s when is_integer(s) and s >= 1 -> s
          s -> raise ArgumentError, "expected an integer >= 1 or a list of integers, got: #{inspect(s)}"
        end

      {pad, lo, hi} = padding(opts[:padding])
      {flat, vec_axes, lead, [len], ch} = Ctx.flatten_batch(tensor, 1)
      type = Nx.type(flat)
      {_, bits} = type
      # 32- and 64-bit elements are moved as they are; narrower types go through their 32-bit form
      {wire, elem} = if bits == 64, do: {type, 8}, else: {if(Nx.Type.float?(type), do: {:f, 32}, else: {:s, 32}), 4}

      case NIF.as_windowed(Ctx.get(), Nx.to_binary(Nx.as_type(flat, wire)), elem, ch, len, window_length, stride, pad, lo, hi) do
        {:ok, out, frames} ->
          out
          |> Nx.from_binary(wire)
          |> Nx.as_type(type)
          |> Nx.reshape(List.to_tuple(lead ++ [frames, window_length]))
          |> Nx.vectorize(vec_axes)

        err ->
          Ctx.raise_nif(err)
      end
    end
  end

  @doc "`NxSignal.overlap_and_add/2` (lib/nx_signal.ex:684-735)."
  def overlap_and_add(tensor, opts \\ []) do
    if Ctx.expr?(tensor) do
      NxSignal.overlap_and_add(tensor, opts)
    else
      opts = Keyword.validate!(opts, [:overlap_length, type: Nx.type(tensor)])
      overlap = opts[:overlap_length]
      {flat, vec_axes, lead, [frames, flen], batch} = Ctx.flatten_batch(tensor, 2)

      if overlap >= flen,
        do: raise(ArgumentError, "overlap_length must be a number less than the window size #{flen}, got: #{inspect(flen)}")

      cplx = Nx.Type.complex?(Nx.type(flat))
      wire = if cplx, do: :c64, else: :f32

      case NIF.overlap_and_add(Ctx.get(), Nx.to_binary(Nx.as_type(flat, wire)), if(cplx, do: 1, else: 0), batch, frames, flen, overlap) do
        {:ok, out, out_len} ->
          out
          |> Nx.from_binary(wire)
          |> Nx.as_type(opts[:type])
          |> Nx.reshape(List.to_tuple(lead ++ [out_len]))
          |> Nx.vectorize(vec_axes)

        err ->
          Ctx.raise_nif(err)
      end
    end
  end

  @doc "`NxSignal.fft_frequencies/2` (lib/nx_signal.ex:154-166)."
  def fft_frequencies(sampling_rate, opts \\ []) do
    opts = Keyword.validate!(opts, [:fft_length, type: {:f, 32}, name: :frequencies, endpoint: false])

    if is_number(sampling_rate) and opts[:endpoint] == false and opts[:type] in [{:f, 32}, :f32] do
      case NIF.fft_frequencies(sampling_rate * 1.0, opts[:fft_length]) do
        {:ok, f} -> f |> Nx.from_binary(:f32) |> Nx.reshape({opts[:fft_length]}, names: [opts[:name]])
        err -> Ctx.raise_nif(err)
      end
    else
      # tensor sampling rates, other types and endpoint: true are O(n) closed forms: the reference computes them
      NxSignal.fft_frequencies(sampling_rate, opts)
    end
  end

  @doc "`NxSignal.mel_filters/4` (lib/nx_signal.ex:397-445): `{mel_bins, fft_length}`."
  def mel_filters(fft_length, mel_bins, sampling_rate, opts \\ []) do
    opts = Keyword.validate!(opts, max_mel: 3016, mel_frequency_spacing: 200 / 3, type: {:f, 32})

    case NIF.mel_filters(fft_length, mel_bins, sampling_rate * 1.0, opts[:max_mel] * 1.0, opts[:mel_frequency_spacing] * 1.0) do
      {:ok, f} ->
        f |> Nx.from_binary(:f32) |> Nx.reshape({mel_bins, fft_length}, names: [:mel, :frequencies]) |> Nx.as_type(opts[:type])

      err ->
        Ctx.raise_nif(err)
    end
  end

  @doc "`NxSignal.stft_to_mel/3` (lib/nx_signal.ex:486-513): `z {frames, frequencies}` (vectorised axes = batch)."
  def stft_to_mel(z, sampling_rate, opts \\ []) do
    if Ctx.expr?(z) do
      NxSignal.stft_to_mel(z, sampling_rate, opts)
    else
      opts = Keyword.validate!(opts, [:fft_length, :mel_bins, :max_mel, :mel_frequency_spacing, type: {:f, 32}])
      nfft = opts[:fft_length] || raise(ArgumentError, "missing :fft_length option")
      mels = opts[:mel_bins] || raise(ArgumentError, "missing :mel_bins option")
      {flat, vec_axes, lead, [frames, zlen], ch} = Ctx.flatten_batch(z, 2)

      case NIF.stft_to_mel(Ctx.get(), Nx.to_binary(Nx.as_type(flat, :c64)), ch, frames, zlen, nfft, mels, sampling_rate * 1.0,
             (opts[:max_mel] || 3016) * 1.0, (opts[:mel_frequency_spacing] || 200 / 3) * 1.0) do
        {:ok, mel} ->
          mel
          |> Nx.from_binary(:f32)
          |> Nx.reshape(List.to_tuple(lead ++ [frames, mels]))
          |> Nx.vectorize(vec_axes)
          |> Nx.rename([:frames, :mel])
          |> Nx.as_type(opts[:type])

        err ->
          Ctx.raise_nif(err)
      end
    end
  end

  @doc """
  `NxSignal.stft/3 |> NxSignal.stft_to_mel/3` in one device call (not a reference head): the spectrum is
  never stored and only the `{frames, mel}` tensor crosses PCIe.  Options: those of `stft/3` plus
  `:mel_bins`, `:max_mel`, `:mel_frequency_spacing`.
  """
  def stft_mel(data, window, opts \\ []) do
    {frame_length} = Nx.shape(window)

    opts =
      Keyword.validate!(opts, [
        :overlap_length,
        :scaling,
        :max_mel,
        :mel_frequency_spacing,
        window_padding: :valid,
        sampling_rate: 100,
        fft_length: :power_of_two,
        mel_bins: 128
      ])

    overlap = opts[:overlap_length] || div(frame_length, 2)
    nfft = fft_len(opts[:fft_length], frame_length)
    {pad, lo, hi} = padding(opts[:window_padding])
    {flat, vec_axes, lead, [len], ch} = Ctx.flatten_batch(data, 1)

    case NIF.stft_mel(Ctx.get(), Nx.to_binary(Nx.as_type(flat, :f32)), ch, len, Nx.to_binary(Nx.as_type(window, :f32)),
           frame_length - overlap, nfft, pad, lo, hi, scaling!(opts[:scaling]), opts[:sampling_rate] * 1.0, opts[:mel_bins],
           (opts[:max_mel] || 3016) * 1.0, (opts[:mel_frequency_spacing] || 200 / 3) * 1.0) do
      {:ok, mel, frames} ->
        mel
        |> Nx.from_binary(:f32)
        |> Nx.reshape(List.to_tuple(lead ++ [frames, opts[:mel_bins]]))
        |> Nx.vectorize(vec_axes)
        |> Nx.rename([:frames, :mel])

      err ->
        Ctx.raise_nif(err)
    end
  end
end

defmodule NxSignalB200.Windows do
  @moduledoc """
  `NxSignal.Windows.*` (lib/nx_signal/windows.ex:33, 57, 98, 160, 225, 278, 341): the same heads, options
  and defaults.  The f32 values are bit-identical to `Nx.BinaryBackend`'s (`nxs_window_f32`); other
  `:type`s are O(n) closed forms and are left to the reference so that they are computed in that type.
  """
  alias NxSignalB200.{Ctx, NIF}

  @kinds %{rectangular: 0, bartlett: 1, triangular: 2, blackman: 3, hamming: 4, hann: 5, kaiser: 6}

  defp gen(kind, n, periodic, beta, eps, opts) do
    type = Nx.Type.normalize!(opts[:type])

    if type == {:f, 32} do
      case NIF.window(@kinds[kind], n, if(periodic, do: 1, else: 0), beta * 1.0, eps * 1.0) do
        {:ok, w} -> w |> Nx.from_binary(:f32) |> Nx.reshape({n}, names: [opts[:name]])
        err -> Ctx.raise_nif(err)
      end
    else
      apply(NxSignal.Windows, kind, [n, opts])
    end
  end

  def rectangular(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, type: :s64)
    # integer ones by default, exactly like the reference (windows.ex:33-36)
    NxSignal.Windows.rectangular(n, opts)
  end

  def bartlett(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, type: {:f, 32})
    gen(:bartlett, n, true, 0.0, 0.0, Keyword.put(opts, :name, nil))
  end

  def triangular(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, [:name, type: {:f, 32}])
    gen(:triangular, n, true, 0.0, 0.0, opts)
  end

  def blackman(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, [:name, is_periodic: true, type: {:f, 32}])
    gen(:blackman, n, opts[:is_periodic], 0.0, 0.0, opts)
  end

  def hamming(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, [:name, is_periodic: true, type: {:f, 32}])
    gen(:hamming, n, opts[:is_periodic], 0.0, 0.0, opts)
  end

  def hann(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, [:name, is_periodic: true, type: {:f, 32}])
    gen(:hann, n, opts[:is_periodic], 0.0, 0.0, opts)
  end

  def kaiser(n, opts \\ []) when is_integer(n) do
    opts = Keyword.validate!(opts, [:name, eps: 1.0e-7, beta: 12.0, is_periodic: true, type: {:f, 32}])
    gen(:kaiser, n, opts[:is_periodic], opts[:beta], opts[:eps], opts)
  end
end

defmodule NxSignalB200.Filters do
  @moduledoc """
  `NxSignal.Filters.firwin/3` (lib/nx_signal/filters.ex:147-279), `median/2` (:17-56), `wiener/2` (:80-110).
  """
  alias NxSignalB200.{Ctx, NIF}

  @windows %{hamming: 4, hann: 5, blackman: 3, bartlett: 1, rectangular: 0}

  def firwin(num_taps, cutoff, opts \\ []) do
    opts = Keyword.validate!(opts, window: :hamming, pass_zero: true, scale: true, sampling_rate: 2.0, type: {:f, 32})

    if not is_list(cutoff), do: raise(ArgumentError, "cutoff must be a list of frequencies, got: #{inspect(cutoff)}")

    nyq = opts[:sampling_rate] / 2.0
    sorted = cutoff |> Enum.map(&(&1 / nyq)) |> Enum.sort()

    if List.first(sorted) <= 0.0,
      do: raise(ArgumentError, "cutoff must be strictly between 0 and Nyquist (exclusive), got: #{List.first(sorted) * nyq}")

    if List.last(sorted) >= 1.0,
      do: raise(ArgumentError, "cutoff must be strictly between 0 and Nyquist (exclusive), got: #{List.last(sorted) * nyq}")

    even_cuts = rem(length(sorted), 2) == 0
    nyquist_gain = (opts[:pass_zero] and even_cuts) or (not opts[:pass_zero] and not even_cuts)

    if nyquist_gain and rem(num_taps, 2) == 0,
      do:
        raise(
          ArgumentError,
          "a filter with non-zero gain at Nyquist (e.g. highpass) requires an odd number of taps, got: #{num_taps}"
        )

    {kind, beta} =
      case opts[:window] do
        {:kaiser, beta} ->
          {6, beta * 1.0}

        w when is_map_key(@windows, w) ->
          {@windows[w], 0.0}

        w ->
          raise ArgumentError,
                "unknown window #{inspect(w)}, supported: :hamming, :hann, :blackman, :bartlett, :rectangular, {:kaiser, beta}"
      end

    if Nx.Type.normalize!(opts[:type]) == {:f, 32} do
      case NIF.firwin(num_taps, Enum.map(cutoff, &(&1 * 1.0)), kind, beta, if(opts[:pass_zero], do: 1, else: 0),
             if(opts[:scale], do: 1, else: 0), opts[:sampling_rate] * 1.0) do
        {:ok, h} -> h |> Nx.from_binary(:f32) |> Nx.reshape({num_taps})
        err -> Ctx.raise_nif(err)
      end
    else
      NxSignal.Filters.firwin(num_taps, cutoff, opts)
    end
  end

  def median(t, opts) do
    if Ctx.expr?(t) or Nx.rank(t) > 3 do
      NxSignal.Filters.median(t, opts)
    else
      opts = Keyword.validate!(opts, [:kernel_shape])

      if Nx.rank(t) != tuple_size(opts[:kernel_shape]),
        do: raise(ArgumentError, "kernel shape must be of the same rank as the tensor")

      case NIF.median(Ctx.get(), Nx.to_binary(Nx.as_type(t, :f32)), Tuple.to_list(Nx.shape(t)), Tuple.to_list(opts[:kernel_shape])) do
        {:ok, out} -> out |> Nx.from_binary(:f32) |> Nx.reshape(Nx.shape(t))
        err -> Ctx.raise_nif(err)
      end
    end
  end

  def wiener(t, opts \\ []) do
    if Ctx.expr?(t) or Nx.rank(t) > 3 do
      NxSignal.Filters.wiener(t, opts)
    else
      opts = Keyword.validate!(opts, noise: nil, kernel_size: 3)
      rank = Nx.rank(t)

      kernel =
        case opts[:kernel_size] do
          k when is_integer(k) -> List.duplicate(k, rank)
          k when is_tuple(k) -> Tuple.to_list(k)
          _ -> raise ArgumentError, "kernel_size must be an integer or tuple"
        end

      type = Nx.type(t)
      f64? = type == {:f, 64}
      wire = if f64?, do: :f64, else: :f32

      case NIF.wiener(Ctx.get(), Nx.to_binary(Nx.as_type(t, wire)), if(f64?, do: 1, else: 0), Tuple.to_list(Nx.shape(t)), kernel,
             if(is_nil(opts[:noise]), do: 0, else: 1), (opts[:noise] || 0.0) * 1.0) do
        # the reference computes in f64 and casts back to the input's type (filters.ex:108-110)
        {:ok, out} -> out |> Nx.from_binary(wire) |> Nx.reshape(Nx.shape(t)) |> Nx.as_type(type)
        err -> Ctx.raise_nif(err)
      end
    end
  end
end

defmodule NxSignalB200.Convolution do
  @moduledoc """
  `NxSignal.Convolution.convolve/3` (lib/nx_signal/convolution.ex:38-58), `correlate/3` (:87-93) and
  `fftconvolve/3` (:252-298).  Both methods give the values of the linear convolution; on the GPU the
  batched FIR form (`{C, L}` with `{1, K}`, or two rank-1 operands) runs the overlap-save kernels and
  other operands of rank <= 3 the direct N-d kernel, whatever `:method` says.  Operands the backend
  does not take (rank > 3 after squeezing, types wider than 32 bits) go to the reference.
  """
  alias NxSignalB200.{Ctx, NIF}

  @mode %{full: 0, same: 1, valid: 2}

  def convolve(in1, in2, opts \\ []) do
    opts = Keyword.validate!(opts, mode: :full, method: :direct)

    if opts[:mode] not in [:full, :same, :valid],
      do: raise(ArgumentError, "expected mode to be one of [:full, :same, :valid], got: #{inspect(opts[:mode])}")

    if opts[:method] not in [:direct, :fft],
      do: raise(ArgumentError, "expected method to be one of [:direct, :fft], got: #{inspect(opts[:method])}")

    dispatch(in1, in2, opts, &NxSignal.Convolution.convolve/3)
  end

  def fftconvolve(in1, in2, opts \\ []) do
    opts = Keyword.validate!(opts, mode: :full, method: :direct)
    if Nx.rank(in1) != Nx.rank(in2), do: raise(ArgumentError, "Rank of in1 and in2 must be equal.")
    dispatch(in1, in2, opts, &NxSignal.Convolution.fftconvolve/3)
  end

  def correlate(in1, in2, opts \\ []) do
    if Ctx.expr?(in1) or Ctx.expr?(in2) do
      NxSignal.Convolution.correlate(in1, in2, opts)
    else
      flipped = Nx.reverse(in2)
      convolve(in1, if(Nx.Type.complex?(Nx.type(in2)), do: Nx.conjugate(flipped), else: flipped), opts)
    end
  end

  defp dispatch(in1, in2, opts, reference) do
    {r1, r2} = {Nx.rank(in1), Nx.rank(in2)}
    type = Nx.Type.merge(Nx.type(in1), Nx.type(in2))
    {_, bits} = type
    wide = bits > 64 or (bits == 64 and not Nx.Type.complex?(type))

    cond do
      Ctx.expr?(in1) or Ctx.expr?(in2) or wide or in1.vectorized_axes != [] or in2.vectorized_axes != [] ->
        reference.(in1, in2, opts)

      r1 == 0 or r2 == 0 ->
        # scalar times tensor: no convolution to run (convolution.ex:98-104 validates the ranks)
        reference.(in1, in2, opts)

      r1 != r2 ->
        raise ArgumentError,
              "NxSignal.convolve/3 requires both inputs to have the same rank or one of them to be a scalar, got #{r1} and #{r2}"

      true ->
        run(in1, in2, type, opts, reference)
    end
  end

  defp run(in1, in2, type, opts, reference) do
    s1 = Tuple.to_list(Nx.shape(in1))
    s2 = Tuple.to_list(Nx.shape(in2))
    cplx = Nx.Type.complex?(type)
    out_type = if cplx, do: {:c, 64}, else: Nx.Type.to_floating(type)
    fir? = not cplx and Enum.drop(s2, -1) |> Enum.all?(&(&1 == 1)) and List.last(s2) >= 16

    cond do
      fir? ->
        # batched FIR form: every leading axis of in1 is a batch axis
        len = List.last(s1)
        ch = div(Enum.product(s1), len)

        case NIF.fir(Ctx.get(), Nx.to_binary(Nx.as_type(in1, :f32)), ch, len, Nx.to_binary(Nx.as_type(Nx.flatten(in2), :f32)),
               @mode[opts[:mode]]) do
          {:ok, y, out_len} ->
            y |> Nx.from_binary(:f32) |> Nx.reshape(List.to_tuple(List.replace_at(s1, -1, out_len))) |> Nx.as_type(out_type)

          err ->
            Ctx.raise_nif(err)
        end

      length(s1) <= 3 ->
        pad3 = fn s -> List.duplicate(1, 3 - length(s)) ++ s end
        wire = if cplx, do: :c64, else: :f32

        case NIF.convolve_nd(Ctx.get(), Nx.to_binary(Nx.as_type(in1, wire)), pad3.(s1), Nx.to_binary(Nx.as_type(in2, wire)),
               pad3.(s2), if(cplx, do: 1, else: 0), @mode[opts[:mode]]) do
          {:ok, out, os} ->
            out |> Nx.from_binary(wire) |> Nx.reshape(List.to_tuple(Enum.drop(os, 3 - length(s1)))) |> Nx.as_type(out_type)

          {:error, :argument_error, _} when opts[:mode] == :valid ->
            raise ArgumentError, "For :valid mode, one must be at least as large as the other in every dimension"

          err ->
            Ctx.raise_nif(err)
        end

      true ->
        reference.(in1, in2, opts)
    end
  end
end

defmodule NxSignalB200.PeakFinding do
  @moduledoc """
  `NxSignal.PeakFinding.argrelmin/2` (lib/nx_signal/peak_finding.ex:131), `argrelmax/2` (:252) and
  `argrelextrema/3` (:340) for the comparators `&Nx.less/2`, `&Nx.greater/2`, `&Nx.less_equal/2`,
  `&Nx.greater_equal/2`; any other comparator function cannot cross a C ABI and runs the reference.
  Returns `%{indices: s64 {n, rank} (-1 padded), valid_indices: count}` like the reference's `nonzero`.
  """
  alias NxSignalB200.{Ctx, NIF}

  def argrelmin(data, opts \\ []), do: run(data, 0, opts, fn -> NxSignal.PeakFinding.argrelmin(data, opts) end)
  def argrelmax(data, opts \\ []), do: run(data, 1, opts, fn -> NxSignal.PeakFinding.argrelmax(data, opts) end)

  def argrelextrema(data, comparator_fn, opts \\ []) do
    cmp =
      cond do
        comparator_fn == (&Nx.less/2) -> 0
        comparator_fn == (&Nx.greater/2) -> 1
        comparator_fn == (&Nx.less_equal/2) -> 2
        comparator_fn == (&Nx.greater_equal/2) -> 3
        true -> nil
      end

    reference = fn -> NxSignal.PeakFinding.argrelextrema(data, comparator_fn, opts) end
    if is_nil(cmp), do: reference.(), else: run(data, cmp, opts, reference)
  end

  defp run(data, cmp, opts, reference) do
    opts = Keyword.validate!(opts, axis: 0, order: 1)

    if Ctx.expr?(data) or data.vectorized_axes != [] or Nx.rank(data) < 1 or Nx.rank(data) > 8 or opts[:order] < 1 do
      reference.()
    else
      shape = Tuple.to_list(Nx.shape(data))
      rank = length(shape)
      axis = if opts[:axis] < 0, do: opts[:axis] + rank, else: opts[:axis]

      case NIF.argrelextrema(Ctx.get(), Nx.to_binary(Nx.as_type(data, :f32)), shape, axis, opts[:order], cmp) do
        {:ok, idx, valid} ->
          %{
            indices: idx |> Nx.from_binary(:s32) |> Nx.reshape({Enum.product(shape), rank}) |> Nx.as_type(:s64),
            valid_indices: Nx.tensor(valid, type: :u64)
          }

        err ->
          Ctx.raise_nif(err)
      end
    end
  end
end
